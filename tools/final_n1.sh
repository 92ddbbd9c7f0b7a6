#!/bin/bash
# Round-end measurements on one GPU: tests, smoke, default bench, reference arm, launch list
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; tail -2 gpurun_out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; tail -1 gpurun_out/r02f_smoke.log
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_refarm.json 2> gpurun_out/r02f_refarm.err; echo "ref rc=$?"
for w in c1 c2 c3; do python bench.py --workload $w --steps 10 --warmup 3 --c5 off > gpurun_out/r02f_bench_$w.json 2>/dev/null; done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02f_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu --c5 off > gpurun_out/r02f_ncu_launches.log 2>&1; echo "ncu list rc=$?"

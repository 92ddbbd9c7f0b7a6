#!/bin/bash
# ncu --set full captures of the dominant kernels (each command has already run to completion without ncu)
cd "$(dirname "$0")/.."
python bench.py --pairs 120000 --steps 1 --warmup 1 --no-cpu --c5 off > gpurun_out/r02d_plain_c4.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:batch_line16_kernel --launch-skip 8 -c 5 -o gpurun_out/r02d_batch_c4 -f python bench.py --pairs 120000 --steps 1 --warmup 1 --no-cpu --c5 off > gpurun_out/r02d_ncu_c4.log 2>&1
REPS=2 python tools/probe_tile.py 100000 8,2 > gpurun_out/r02d_plain_c3.log 2>&1 || exit 1
REPS=2 ncu --set full --clock-control none --import-source on -k regex:tile_fill_kernel --launch-skip 1 -c 1 -o gpurun_out/r02d_tile_c3_nw -f python tools/probe_tile.py 100000 8,2 > gpurun_out/r02d_ncu_c3nw.log 2>&1
MODE=1 REPS=2 ncu --set full --clock-control none --import-source on -k regex:tile_fill_kernel --launch-skip 1 -c 1 -o gpurun_out/r02d_tile_c3_sw -f python tools/probe_tile.py 100000 8,2 > gpurun_out/r02d_ncu_c3sw.log 2>&1
ls -la gpurun_out/r02d_*.ncu-rep
for n in r02d_batch_c4 r02d_tile_c3_nw r02d_tile_c3_sw; do python tools/summarize_ncu.py gpurun_out/$n.ncu-rep > gpurun_out/${n}_ncu.txt 2>/dev/null; done
rm -f gpurun_out/r02d_tile_c3_nw.ncu-rep gpurun_out/r02d_batch_c4.ncu-rep
ls -la gpurun_out/ | head -20

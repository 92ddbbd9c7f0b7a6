#!/bin/bash
# Round-end measurements on one 8-GPU box: the bench line (config 4 strong scaling + config 5) at N = 8, 4, 2
cd "$(dirname "$0")/.."
for N in ${NS:-4 2}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu 2> gpurun_out/r02f_bench_n$N.err | grep '^{' > gpurun_out/r02f_bench_n$N.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r02f_bench_n$N.json")); c=d.get("c5") or {}
print("N=$N value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]),"e2e ms",round(d["e2e"]["ms_per_step"],3),"verified",d["verified"]["mismatches"],"| c5",c.get("seconds"),c.get("value"),[round(x,1) for x in c.get("fill_ms_per_rank",[])],c.get("checks",{}).get("score_equals_cpu_oracle"))
PY
done
nvidia-smi topo -m > gpurun_out/r02f_topo_n8.txt 2>&1

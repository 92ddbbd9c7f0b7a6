#!/bin/bash
# Dev tool: config 5 at N = 1 / 2 on one box, kernel and residency variants
cd "$(dirname "$0")/.."
run() { # N tag env...
  N=$1; tag=$2; shift 2
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench_c5.py --length 1000000 --steps 2 2> gpurun_out/s2_c5_n${N}_$tag.err | grep '^{' > gpurun_out/s2_c5_n${N}_$tag.json
  python - <<PY
import json
d=json.load(open("gpurun_out/s2_c5_n${N}_$tag.json"))
print("N=$N $tag", round(d["seconds"],4), "s", round(d["value"]), "GCUPS fills", [round(x,1) for x in d["fill_ms_per_rank"]], d["checks"]["score_equals_cpu_oracle"], d["checks"]["rescore_equals_score"])
PY
}
run 2 tile2 SA_TILE=8,2
run 2 tile3 SA_TILE=8,2 SA_LONG_BLOCKS_PER_SM=3
run 2 tile1 SA_TILE=8,2 SA_LONG_BLOCKS_PER_SM=1
run 1 ckpt2 X=1
run 1 ckpt3 SA_LONG_BLOCKS_PER_SM=3
python tools/probe_tile.py 500000 8,2 2>&1 | tail -1
python tools/probe_tile.py 100000 8,2 2>&1 | tail -1

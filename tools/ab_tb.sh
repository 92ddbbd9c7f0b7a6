#!/bin/bash
# A/B on config 4, one GPU: threads per traceback block next to the fill blocks of the following chunk
cd "$(dirname "$0")/.."
show='import json,sys;d=json.loads(sys.stdin.read());print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e ms",round(d["e2e"]["ms_per_step"],2),"fill ms",round(d["roofline"]["kernel_ms_per_step"],3),"tb",round(d["roofline"]["traceback_ms_per_step"],2),"launches",d["gpu_launches"])'
run() { echo "== $*"; env "$@" python bench.py --no-cpu --c5 off --verify-pairs 0 2>/dev/null | python -c "$show"; }
run SA_TB_THREADS=64
run SA_TB_THREADS=32
run SA_TB_THREADS=64 SA_TB_BLOCKS_PER_SM=2
run SA_TB_THREADS=32 SA_TB_BLOCKS_PER_SM=2

#!/bin/bash
# A/B of the chunk sizes of the device-resident and of the host-buffer batch (config 4, one GPU)
cd "$(dirname "$0")/.."
show='import json,sys;d=json.loads(sys.stdin.read());print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e ms",round(d["e2e"]["ms_per_step"],2),"fill ms",round(d["roofline"]["kernel_ms_per_step"],3),"tb",round(d["roofline"]["traceback_ms_per_step"],2),"launches",d["gpu_launches"])'
run() { echo "== $*"; env "$@" python bench.py --no-cpu --c5 off --verify-pairs 0 2>/dev/null | python -c "$show"; }
run SA_X=default
run SA_DEV_DIRS_BUDGET_MB=2500 SA_HOST_SCHEDULE=1,2,3,3,3,3,3,3,3,3,2,2,1
run SA_DEV_DIRS_BUDGET_MB=3000 SA_HOST_SCHEDULE=1,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,1
run SA_DEV_DIRS_BUDGET_MB=2000 SA_HOST_SCHEDULE=1,2,3,4,4,4,4,4,3,2,1

#!/bin/bash
# A/B of chunk schedules / traceback residency of the device-resident and of the host-buffer batch (config 4, one GPU)
cd "$(dirname "$0")/.."
show='import json,sys;d=json.loads(sys.stdin.read());print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e ms",round(d["e2e"]["ms_per_step"],2),"fill ms",round(d["roofline"]["kernel_ms_per_step"],3),"tb",round(d["roofline"]["traceback_ms_per_step"],2),"launches",d["gpu_launches"])'
run() { echo "== $*"; env "$@" python bench.py --no-cpu --c5 off --verify-pairs 0 2>/dev/null | python -c "$show"; }
run SA_X=default
run SA_HOST_SCHEDULE=1,1,2,4,8,8,8,8,8,8,4,2,1,1
run SA_HOST_SCHEDULE=1,2,4,6,8,8,8,8,6,4,2,1 SA_TB_BLOCKS_PER_SM=2
run SA_HOST_SCHEDULE=2,4,6,8,8,8,8,8,6,4,2

#!/bin/bash
# Dev tool: one config-5 slice (125 000 columns x 950 k rows) under different residency caps (SA_LONG_BLOCKS_PER_SM)
cd "$(dirname "$0")/.."
for cfg in "SA_TILE=8,2 1" "SA_TILE=8,2 2" "SA_TILE=8,2 3" "SA_TILE=8,2 4" "SA_TILE=8,2 8" "SA_LONG_R=16 1" "SA_LONG_R=16 2" "SA_LONG_R=8 1" "SA_LONG_R=8 2"; do
  set -- $cfg
  echo "== $1 blocks/SM $2"
  env $1 SA_LONG_BLOCKS_PER_SM=$2 SLICE=${SLICE:-125000} timeout 300 python tools/probe_strip_times.py 1000000 2>&1 | grep -E "kernel span|strip     0:|end-to-end lag|strips=" | head -6
done

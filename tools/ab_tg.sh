cd /root/repo
export SA_B200_LIB=/root/repo/sequence-alignment-gpu_b200/libsa_b200_few.so
echo "== linked kernel without the per-step stamp (strip API, SA_LONG_DBG -> linked variant)"
SA_TILE=8,2 python tools/probe_strip_times.py 100000 2>&1 | grep -E "kernel span|lag per strip|top-row waits|end-to-end"
SLICE=125000 SA_TILE=8,2 python tools/probe_strip_times.py 1000000 2>&1 | grep -E "kernel span|lag per strip|top-row waits|end-to-end"
export SA_B200_LIB=/root/repo/sequence-alignment-gpu_b200/libsa_b200_dbgp.so
echo "== plain kernel with stamps"
SA_LONG_DBG_PLAIN=1 SA_TILE=8,2 python tools/probe_strip_times.py 100000 2>&1 | grep -E "kernel span|strip     [0123]:|lag per strip|top-row waits|end-to-end"
SLICE=125000 SA_LONG_DBG_PLAIN=1 SA_TILE=8,2 python tools/probe_strip_times.py 1000000 2>&1 | grep -E "kernel span|lag per strip|top-row waits|end-to-end"
python tools/probe_tile.py 100000 8,2 | tail -1

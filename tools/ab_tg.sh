cd /root/repo
export SA_B200_LIB=/root/repo/sequence-alignment-gpu_b200/libsa_b200_dbgp.so
SA_LONG_DBG_PLAIN=1 SA_TILE=8,2 python tools/probe_strip_times.py 100000 2>&1 | grep -E "kernel span|lag per strip|top-row waits|end-to-end|stalls after|gaps between"

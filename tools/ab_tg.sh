cd /root/repo
for tag in "" _tg4 _tg4r2 _tg16; do
  export SA_B200_LIB=/root/repo/sequence-alignment-gpu_b200/libsa_b200$tag.so
  echo "== lib$tag"
  python tools/probe_tile.py 100000 8,2 4,4 2>&1 | tail -2
  SLICE=125000 REPS=3 python tools/probe_tile.py 1000000 8,2 2>&1 | tail -1
  SLICE=250000 REPS=3 python tools/probe_tile.py 1000000 8,2 2>&1 | tail -1
done
export SA_B200_LIB=/root/repo/sequence-alignment-gpu_b200/libsa_b200_tg4.so
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile" 2>&1 | tail -3

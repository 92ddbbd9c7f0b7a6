cd /root/repo
export SA_B200_LIB=/root/repo/sequence-alignment-gpu_b200/libsa_b200_sw.so
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile or random_pairs or ties or goldens or traceback_variants or gap_penalty or known_answer or identity" 2>&1 | tail -3
MODE=1 python tools/probe_tile.py 100000 8,2 4,4 strip:8 2>&1 | tail -3
MODE=1 PROTEIN=1 python tools/probe_tile.py 5000 8,2 4,4 strip:4 2>&1 | tail -3
python tools/probe_tile.py 100000 8,2 2>&1 | tail -1

cd /root/repo
python tools/probe_tile.py 100000 8,2 4,4 2>&1 | tail -2
MODE=1 python tools/probe_tile.py 100000 8,2 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2

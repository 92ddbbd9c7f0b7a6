#!/bin/bash
# Dev tool: tiled vs one-column kernel on the big shapes of config 5 (one GPU)
cd "$(dirname "$0")/.."
REPS=2 python tools/probe_tile.py 500000 8,2 strip:16 2>&1 | tail -2
SLICE=500000 REPS=2 python tools/probe_tile.py 1000000 8,2 strip:16 2>&1 | tail -2
M=390000 REPS=2 python tools/probe_tile.py 1000000 8,2 strip:12 strip:16 2>&1 | tail -3
M=65536 REPS=2 python tools/probe_tile.py 1000000 8,2 strip:8 2>&1 | tail -2

#!/bin/bash
# ncu --set full capture of the batch traceback kernel (the command has already run to completion without ncu):
# launch 0 of the capture = a chunk's traceback with one block per SM (the shape that runs beside the next fill),
# launch 1 = the last chunk's traceback with the full grid
cd "$(dirname "$0")/.."
T=${TAG:-r02d}
python bench.py --pairs 180000 --steps 1 --warmup 1 --no-cpu --c5 off > gpurun_out/${T}_plain_tb.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:batch_traceback_kernel --launch-skip 2 -c 2 -o gpurun_out/${T}_batch_tb -f python bench.py --pairs 180000 --steps 1 --warmup 1 --no-cpu --c5 off > gpurun_out/${T}_ncu_tb.log 2>&1
python tools/summarize_ncu.py gpurun_out/${T}_batch_tb.ncu-rep > gpurun_out/${T}_batch_tb_ncu.txt 2>/dev/null
rm -f gpurun_out/${T}_batch_tb.ncu-rep
ls -la gpurun_out/${T}_*tb*

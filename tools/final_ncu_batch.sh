#!/bin/bash
# ncu --set full capture of the packed batch kernel (the command has already run to completion without ncu)
cd "$(dirname "$0")/.."
T=${TAG:-r02d}
python bench.py --pairs 120000 --steps 1 --warmup 1 --no-cpu --c5 off > gpurun_out/${T}_plain_c4.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:batch_line16_kernel --launch-skip 8 -c 5 -o gpurun_out/${T}_batch_c4 -f python bench.py --pairs 120000 --steps 1 --warmup 1 --no-cpu --c5 off > gpurun_out/${T}_ncu_c4.log 2>&1
python tools/summarize_ncu.py gpurun_out/${T}_batch_c4.ncu-rep > gpurun_out/${T}_batch_c4_ncu.txt 2>/dev/null
ncu -i gpurun_out/${T}_batch_c4.ncu-rep --page source --csv > gpurun_out/${T}_batch_c4_source.csv 2>/dev/null
rm -f gpurun_out/${T}_batch_c4.ncu-rep
ls -la gpurun_out/${T}_*

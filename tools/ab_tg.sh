cd /root/repo
ncu --set full --clock-control none --import-source on -k regex:batch_line16_kernel --launch-skip 10 -c 1 -o gpurun_out/r02c_batch_c4 -f python bench.py --pairs 120000 --steps 1 --warmup 1 --no-cpu --c5 off > gpurun_out/r02c_ncu_c4.log 2>&1
ncu -i gpurun_out/r02c_batch_c4.ncu-rep --page source --csv > gpurun_out/r02c_batch_c4_source.csv 2>/dev/null
python tools/summarize_ncu.py gpurun_out/r02c_batch_c4.ncu-rep > gpurun_out/r02c_batch_c4_ncu.txt 2>/dev/null
rm -f gpurun_out/r02c_batch_c4.ncu-rep
ls -la gpurun_out | head

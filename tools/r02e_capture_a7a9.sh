# ncu --set full of the a-7 threshold kernel (80-row canvases) and the bulk-copy greedy CTC kernel (64 and 1024 lines)
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'SrcGrayBatch' -s 3 -c 2 -f -o /tmp/r02e_a7 python tools/stages.py > gpurun_out/r02e_ncu_a7.log 2>&1
ncu -i /tmp/r02e_a7.ncu-rep --page raw --csv > gpurun_out/r02e_a7_full_raw.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'ctc_greedy_bulk' -s 13 -c 6 -f -o /tmp/r02e_a9 python tools/stages.py > gpurun_out/r02e_ncu_a9.log 2>&1
ncu -i /tmp/r02e_a9.ncu-rep --page raw --csv > gpurun_out/r02e_a9_full_raw.csv 2>/dev/null
wc -l gpurun_out/r02e_a7_full_raw.csv gpurun_out/r02e_a9_full_raw.csv; tail -3 gpurun_out/r02e_ncu_a7.log

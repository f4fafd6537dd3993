# ncu evidence for the recognition path (run each command plain first, then under ncu)
python tools/layer_profile.py ocr 512 > gpurun_out/layers_ocr512.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 16 -c 16 -f -o gpurun_out/r01d_ocr_conv python tools/layer_profile.py ocr 512 > gpurun_out/ncu_ocr_conv.log 2>&1
ncu -i gpurun_out/r01d_ocr_conv.ncu-rep --page raw --csv > gpurun_out/r01d_ocr_conv_full_raw.csv 2>/dev/null
python tools/pack_bench.py > gpurun_out/pack_bench.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'ctc_greedy|ocr_canvas' -s 4 -c 4 -f -o gpurun_out/r01d_a7a9 python tools/pack_bench.py > gpurun_out/ncu_a7a9.log 2>&1
ncu -i gpurun_out/r01d_a7a9.ncu-rep --page raw --csv > gpurun_out/r01d_a7a9_full_raw.csv 2>/dev/null
tail -4 gpurun_out/layers_ocr512.log; wc -l gpurun_out/r01d_ocr_conv_full_raw.csv gpurun_out/r01d_a7a9_full_raw.csv

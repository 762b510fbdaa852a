set -x
ncu --set full --clock-control none --import-source on -k regex:score_select_stream -s 3 -c 1 -o gpurun_out/r02_ncu_stream_c4_v2 -f python tools/ncu_small_kernels.py c4 > gpurun_out/r02_ncu_stream_c4_v2.log 2>&1
ncu -i gpurun_out/r02_ncu_stream_c4_v2.ncu-rep --page raw --csv > gpurun_out/r02_ncu_stream_c4_v2_raw.csv
ncu --set full --clock-control none --import-source on -k regex:score_select_tc.*Reservoir -s 1 -c 1 -o gpurun_out/r02_ncu_c5_k100 -f python tools/launch_list_driver.py c5 > gpurun_out/r02_ncu_c5_k100.log 2>&1
ncu -i gpurun_out/r02_ncu_c5_k100.ncu-rep --page raw --csv > gpurun_out/r02_ncu_c5_k100_raw.csv
tail -2 gpurun_out/r02_ncu_stream_c4_v2.log gpurun_out/r02_ncu_c5_k100.log
ls -la gpurun_out/*.csv

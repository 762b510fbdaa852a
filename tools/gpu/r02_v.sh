set -x
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:SelectReservoir -s 2 -c 1 -o gpurun_out/r02_ncu_c5_k100 -f python tools/launch_list_driver.py c5 > gpurun_out/r02_ncu_c5_k100.log 2>&1
ncu -i gpurun_out/r02_ncu_c5_k100.ncu-rep --page raw --csv > gpurun_out/r02_ncu_c5_k100_raw.csv
rm -f gpurun_out/*.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_refscale_f32.csv python tools/launch_list_driver.py refscale > gpurun_out/refscale.log 2>&1
tail -n 2 gpurun_out/refscale.log

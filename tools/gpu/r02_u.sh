set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_1gpu.log
tail -n 3 gpurun_out/r02_pytest_gpu_1gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r02_bench_final.err
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:SelectReservoir -s 1 -c 1 -o gpurun_out/r02_ncu_c5_k100 -f python tools/launch_list_driver.py c5 > gpurun_out/r02_ncu_c5_k100.log 2>&1
ncu -i gpurun_out/r02_ncu_c5_k100.ncu-rep --page raw --csv > gpurun_out/r02_ncu_c5_k100_raw.csv
ls -la gpurun_out/r02_ncu_c5_k100_raw.csv
rm -f gpurun_out/*.ncu-rep

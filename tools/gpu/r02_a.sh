set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_a.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
ncu --set full --clock-control none --import-source on -k regex:score_select_tc -s 3 -c 1 -o gpurun_out/r02_ncu_c3_a -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_c3_a.log 2>&1
ncu -i gpurun_out/r02_ncu_c3_a.ncu-rep --page raw --csv > gpurun_out/r02_ncu_c3_a_raw.csv 2>/dev/null
tail -3 gpurun_out/r02_pytest_a.log; cat gpurun_out/r02_bench_a.json

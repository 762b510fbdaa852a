set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_f.log
python tools/bench_large_k.py > gpurun_out/r02_large_k_f.jsonl 2> gpurun_out/r02_large_k_f.err
for w in c2 c2cos; do
ncu --set full --clock-control none --import-source on -k regex:score_select_tc -s 2 -c 1 -o gpurun_out/r02_ncu_${w}_tier1 -f python tools/launch_list_driver.py $w > gpurun_out/r02_ncu_${w}_tier1.log 2>&1
ncu -i gpurun_out/r02_ncu_${w}_tier1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_${w}_tier1_raw.csv
done
tail -12 gpurun_out/r02_pytest_f.log; cat gpurun_out/r02_large_k_f.jsonl

set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_normalize_bits.py -m gpu -q -x > gpurun_out/r02_pytest_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_m.log
python tools/latency_probe.py bf16 > gpurun_out/r02_latency_probe3_bf16.jsonl 2> gpurun_out/r02_latency_probe3.err
ncu --set full --clock-control none --import-source on -k regex:score_select_stream -s 3 -c 1 -o gpurun_out/r02_ncu_stream_c4 -f python tools/ncu_small_kernels.py c4 > gpurun_out/r02_ncu_stream_c4.log 2>&1
ncu -i gpurun_out/r02_ncu_stream_c4.ncu-rep --page raw --csv > gpurun_out/r02_ncu_stream_c4_raw.csv
ncu -i gpurun_out/r02_ncu_stream_c4.ncu-rep --page source --csv > gpurun_out/r02_ncu_stream_c4_source.csv
tail -3 gpurun_out/r02_pytest_m.log; cut -c1-330 gpurun_out/r02_latency_probe3_bf16.jsonl

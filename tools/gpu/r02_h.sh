set -x
python -m pytest tests/test_gpu_tier1.py tests/test_gpu_parity.py tests/test_gpu_large_k.py -m gpu -q -x > gpurun_out/r02_pytest_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_h.log
python tools/bench_configs.py c2 > gpurun_out/r02_c2_h.jsonl 2> gpurun_out/r02_c2_h.err
tail -5 gpurun_out/r02_pytest_h.log; cat gpurun_out/r02_c2_h.jsonl | cut -c1-700

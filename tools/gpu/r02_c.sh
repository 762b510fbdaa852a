set -x
python -m pytest tests/test_gpu_normalize_bits.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_pytest_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_c.log
python tools/bench_configs.py ingest > gpurun_out/r02_ingest_c.jsonl 2> gpurun_out/r02_ingest_c.err
tail -8 gpurun_out/r02_pytest_c.log; cat gpurun_out/r02_ingest_c.jsonl

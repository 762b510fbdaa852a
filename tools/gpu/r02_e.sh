set -x
python -m pytest tests/test_gpu_normalize_bits.py -m gpu -q -x > gpurun_out/r02_pytest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_e.log
python tools/bench_configs.py ingest > gpurun_out/r02_ingest_e.jsonl 2> gpurun_out/r02_ingest_e.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err
tail -4 gpurun_out/r02_pytest_e.log; cat gpurun_out/r02_ingest_e.jsonl; tail -5 gpurun_out/r02_bench_e.err; cat gpurun_out/r02_bench_e.json

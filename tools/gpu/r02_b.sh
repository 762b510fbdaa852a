set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_b.log
python tools/bench_configs.py ingest > gpurun_out/r02_ingest_b.jsonl 2> gpurun_out/r02_ingest_b.err
tail -25 gpurun_out/r02_pytest_b.log; cat gpurun_out/r02_ingest_b.jsonl

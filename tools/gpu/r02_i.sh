set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_i.log
python tools/bench_configs.py c2 > gpurun_out/r02_c2_i.jsonl 2> gpurun_out/r02_c2_i.err
tail -30 gpurun_out/r02_pytest_i.log | cut -c1-300; cat gpurun_out/r02_c2_i.jsonl | cut -c1-800

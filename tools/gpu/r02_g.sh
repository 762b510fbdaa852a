set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_g.log
python tools/bench_configs.py c2 > gpurun_out/r02_c2_g.jsonl 2> gpurun_out/r02_c2_g.err
tail -12 gpurun_out/r02_pytest_g.log; cat gpurun_out/r02_c2_g.jsonl; tail -3 gpurun_out/r02_c2_g.err

set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_l.log
python tools/latency_probe.py bf16 > gpurun_out/r02_latency_probe2_bf16.jsonl 2> gpurun_out/r02_latency_probe2.err
python tools/latency_probe.py f32 > gpurun_out/r02_latency_probe2_f32.jsonl 2>> gpurun_out/r02_latency_probe2.err
tail -4 gpurun_out/r02_pytest_l.log; cat gpurun_out/r02_latency_probe2_bf16.jsonl gpurun_out/r02_latency_probe2_f32.jsonl | cut -c1-330; tail -3 gpurun_out/r02_latency_probe2.err

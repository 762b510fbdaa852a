set -x
python -m pytest tests/test_gpu_tier1.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_pytest_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_k.log
python tools/latency_probe.py bf16 > gpurun_out/r02_latency_probe_bf16.jsonl 2> gpurun_out/r02_latency_probe.err
python tools/latency_probe.py f32 > gpurun_out/r02_latency_probe_f32.jsonl 2>> gpurun_out/r02_latency_probe.err
tail -4 gpurun_out/r02_pytest_k.log; cat gpurun_out/r02_latency_probe_bf16.jsonl gpurun_out/r02_latency_probe_f32.jsonl; tail -3 gpurun_out/r02_latency_probe.err

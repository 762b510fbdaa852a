set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_normalize_bits.py tests/test_gpu_large_k.py -m gpu -q -x > gpurun_out/r02_pytest_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_n.log
python tools/latency_probe.py bf16 > gpurun_out/r02_latency_probe4_bf16.jsonl 2> gpurun_out/r02_latency_probe4.err
python tools/latency_probe.py f32 > gpurun_out/r02_latency_probe4_f32.jsonl 2>> gpurun_out/r02_latency_probe4.err
tail -3 gpurun_out/r02_pytest_n.log; cut -c1-330 gpurun_out/r02_latency_probe4_bf16.jsonl gpurun_out/r02_latency_probe4_f32.jsonl

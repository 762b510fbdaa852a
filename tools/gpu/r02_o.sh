set -x
python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_gpu_partial_rows.py tests/test_gpu_large_k.py -m gpu -q -x > gpurun_out/r02_pytest_o_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_o_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_o_n2.json 2> gpurun_out/r02_bench_o_n2.err
tail -5 gpurun_out/r02_pytest_o_2gpu.log; tail -5 gpurun_out/r02_bench_o_n2.err; cat gpurun_out/r02_bench_o_n2.json | cut -c1-3000

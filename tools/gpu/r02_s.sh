set -x
python -m pytest tests/test_gpu_multi.py tests/test_gpu_partial_rows.py tests/test_gpu_host_pipeline.py -q -m gpu > gpurun_out/r02_pytest_8gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_8gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r02_bench_n8.json
tail -5 gpurun_out/r02_bench_n8.err

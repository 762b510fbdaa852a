set -x
python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_2gpu.log
tail -n 3 gpurun_out/r02_pytest_gpu_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
echo "bench rc=$?"
tail -c 800 gpurun_out/r02_bench_n2.json

set -x
RDB_BENCH_STEPS=5 python tools/bench_single_process_multi_gpu.py > gpurun_out/r02_sp_n8_pull.jsonl 2> gpurun_out/sp.err
cat gpurun_out/r02_sp_n8_pull.jsonl; tail -3 gpurun_out/sp.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_n8_pull.json 2> gpurun_out/r02_bench_n8.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_n8_pull.json

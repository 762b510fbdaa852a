set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c2_l2.csv python tools/launch_list_driver.py c2 > gpurun_out/r02_launches_c2_l2.log 2>&1
AB_ROUNDS=4 python tools/ab_knobs.py 1250000 768 65536 10 bf16 IP 'tc_chunks=0' 'tc_chunks=2' 'tc_chunks=4' 'tc_chunks=7' 'tc_chunks=13' 'tc_chunks=26' > gpurun_out/r02_ab_chunks_shard.jsonl 2> gpurun_out/r02_ab_chunks_shard.err
cat gpurun_out/r02_ab_chunks_shard.jsonl; tail -3 gpurun_out/r02_ab_chunks_shard.err

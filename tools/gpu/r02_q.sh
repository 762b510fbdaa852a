set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_q.json 2> gpurun_out/r02_bench_q.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_q_ref.json 2>> gpurun_out/r02_bench_q.err
ncu --set full --clock-control none --import-source on -k regex:ingest_fast -s 2 -c 1 -o gpurun_out/r02_ncu_ingest_fast -f python tools/ncu_ingest.py > gpurun_out/r02_ncu_ingest_fast.log 2>&1
ncu -i gpurun_out/r02_ncu_ingest_fast.ncu-rep --page raw --csv > gpurun_out/r02_ncu_ingest_fast_raw.csv
ncu --set full --clock-control none --import-source on -k regex:score_select_tc -s 2 -c 1 -o gpurun_out/r02_ncu_c2_l2_normslice -f python tools/launch_list_driver.py c2 > gpurun_out/r02_ncu_c2_l2_normslice.log 2>&1
ncu -i gpurun_out/r02_ncu_c2_l2_normslice.ncu-rep --page raw --csv > gpurun_out/r02_ncu_c2_l2_normslice_raw.csv
ncu --set full --clock-control none --import-source on -k regex:score_select_tc.*Reservoir -s 1 -c 1 -o gpurun_out/r02_ncu_c5_k100 -f python tools/launch_list_driver.py c5 > gpurun_out/r02_ncu_c5_k100.log 2>&1
ncu -i gpurun_out/r02_ncu_c5_k100.ncu-rep --page raw --csv > gpurun_out/r02_ncu_c5_k100_raw.csv
AB_ROUNDS=5 python tools/ab_knobs.py 1000000 768 256 15 bf16 IP 'tc_chunks=0' > gpurun_out/r02_mid256_staged_merge.jsonl 2>> gpurun_out/r02_bench_q.err
cat gpurun_out/r02_bench_q.json | cut -c1-200; tail -3 gpurun_out/r02_bench_q.err; cat gpurun_out/r02_mid256_staged_merge.jsonl | cut -c1-300

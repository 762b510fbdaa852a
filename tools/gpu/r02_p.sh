set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_mid256.csv python tools/launch_list_driver.py mid256 > gpurun_out/r02_launches_mid256.log 2>&1
AB_ROUNDS=6 python tools/ab_knobs.py 1000000 768 256 15 bf16 IP 'tc_chunks=0' 'tc_cta_group=1' 'tc_chunks=37' 'tc_chunks=111' 'tc_lockstep=0' > gpurun_out/r02_ab_mid256.jsonl 2> gpurun_out/r02_ab_mid256.err
AB_ROUNDS=6 python tools/ab_knobs.py 1000000 768 512 15 bf16 IP 'tc_chunks=0' 'tc_cta_group=1' > gpurun_out/r02_ab_mid512.jsonl 2>> gpurun_out/r02_ab_mid256.err
AB_ROUNDS=6 python tools/ab_knobs.py 1000000 768 128 15 bf16 IP 'tc_chunks=0' 'tc_chunks=74' > gpurun_out/r02_ab_mid128.jsonl 2>> gpurun_out/r02_ab_mid256.err
cat gpurun_out/r02_ab_mid256.jsonl gpurun_out/r02_ab_mid512.jsonl gpurun_out/r02_ab_mid128.jsonl | cut -c1-420

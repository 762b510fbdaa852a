set -x
ncu --set full --clock-control none --import-source on -k regex:ingest_rows -s 2 -c 1 -o gpurun_out/r02_ncu_ingest_norm -f python tools/ncu_ingest.py > gpurun_out/r02_ncu_ingest_norm.log 2>&1
ncu -i gpurun_out/r02_ncu_ingest_norm.ncu-rep --page raw --csv > gpurun_out/r02_ncu_ingest_norm_raw.csv
ncu -i gpurun_out/r02_ncu_ingest_norm.ncu-rep --page source --csv > gpurun_out/r02_ncu_ingest_norm_source.csv
tail -3 gpurun_out/r02_ncu_ingest_norm.log

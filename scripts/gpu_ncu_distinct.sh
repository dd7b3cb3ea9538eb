#!/bin/bash
# ncu full capture of the distinct-key IMMA kernels at N = 677 (262144 rows)
python scripts/profile_distinct.py 262144 1 hps677 > gpurun_out/plain_distinct.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:imma -c 2 -f -o gpurun_out/r2_imma_N677 python scripts/profile_distinct.py 262144 1 hps677 > gpurun_out/ncu_distinct.log 2>&1
echo rc=$?
tail -n 3 gpurun_out/ncu_distinct.log

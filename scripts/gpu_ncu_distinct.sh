#!/bin/bash
# ncu full capture of the distinct-key IMMA kernels (262144 rows): gpu_ncu_distinct.sh [cfg ...]   (default hps677)
for cfg in ${@:-hps677}; do
  N=${cfg//[!0-9]/}
  python scripts/profile_distinct.py 262144 1 $cfg > gpurun_out/plain_distinct.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:imma -c 2 -f -o gpurun_out/r2_imma_N$N python scripts/profile_distinct.py 262144 1 $cfg > gpurun_out/ncu_distinct.log 2>&1
  echo $cfg rc=$?
  tail -n 2 gpurun_out/ncu_distinct.log
done

#!/bin/bash
# ncu full capture of the merged-accumulator IMMA kernels at N = 677 (262144 rows), variant in $1
export NTRU_B200_LIB=$PWD/experiments/ab/libntru_b200_$1.so
python scripts/profile_distinct.py 262144 1 hps677 > gpurun_out/plain_distinct.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:imma -c 2 -f -o gpurun_out/r2_imma_N677_$1 python scripts/profile_distinct.py 262144 1 hps677 > gpurun_out/ncu_distinct.log 2>&1
echo rc=$?
tail -n 2 gpurun_out/ncu_distinct.log

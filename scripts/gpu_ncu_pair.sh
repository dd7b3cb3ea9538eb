#!/bin/bash
# ncu evidence at HEAD (lo + hi schedule): full captures of the three tcgen05 kernels at 1 M rows (N = 509, 821), launch list of bench.py
set -o pipefail
python scripts/profile_target.py 1000000 1 0 hps509 > gpurun_out/plain509.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_umma_pair -c 3 -f -o gpurun_out/r2c_pair_N509 python scripts/profile_target.py 1000000 1 0 hps509 > gpurun_out/ncu509.log 2>&1
echo ncu509 rc=$?
python scripts/profile_target.py 1000000 1 0 hps821 > gpurun_out/plain821.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_umma_pair -c 3 -f -o gpurun_out/r2c_pair_N821 python scripts/profile_target.py 1000000 1 0 hps821 > gpurun_out/ncu821.log 2>&1
echo ncu821 rc=$?
python bench.py --steps 2 --warmup 3 --no-cpu --configs '' > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_ncu_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu --configs '' > gpurun_out/ncu_bench.log 2>&1
echo launchlist rc=$?
for f in gpurun_out/ncu509.log gpurun_out/ncu821.log gpurun_out/ncu_bench.log; do tail -n 3 $f; done
ls -la gpurun_out/*.ncu-rep

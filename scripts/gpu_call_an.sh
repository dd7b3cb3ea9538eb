#!/bin/bash
# merged-accumulator IMMA product: parity of the IMMA path first, then A/B of the register / step-group variants
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q -k "distinct or general or schedules or golden or random or keys or small or ragged or string or class" 2>&1 | tail -8 | tee gpurun_out/r2_imma_merged_tests.txt
for v in head s2b6 s3b6 s2b5 s3b5 s6b4 s3b4; do
  echo "== $v" | tee -a gpurun_out/r2_imma_merged_ab.txt
  NTRU_B200_LIB=$PWD/experiments/ab/libntru_b200_$v.so timeout 300 python scripts/bench_configs.py c3 c3_509 c3_821 c3_167 2>&1 | tail -4 | tee -a gpurun_out/r2_imma_merged_ab.txt
done

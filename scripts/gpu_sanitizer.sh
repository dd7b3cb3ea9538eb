#!/bin/bash
# one compute-sanitizer tool per call ($1 = racecheck | initcheck | memcheck | synccheck)
tool=$1
timeout 300 python scripts/sanitize_target.py 701 8192 > gpurun_out/sanitize_plain.log 2>&1; echo plain rc=$?; tail -2 gpurun_out/sanitize_plain.log
timeout 1200 compute-sanitizer --tool $tool --log-file gpurun_out/r2_sanitizer_${tool}_N701.log python scripts/sanitize_target.py 701 8192 > gpurun_out/sanitize_${tool}.out 2>&1; echo $tool rc=$?
tail -3 gpurun_out/sanitize_${tool}.out; tail -15 gpurun_out/r2_sanitizer_${tool}_N701.log | cut -c1-300

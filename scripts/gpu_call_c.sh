#!/bin/bash
# IMMA B fragments in registers: distinct-key parity + config 3 timing, sum at 10 M and 1.25 M rows
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "distinct or verify_keys or keygen or outside_baseline or golden or random_batch" 2>&1 | tail -5
timeout 600 python scripts/bench_configs.py c3 c3_509 c3_821 c3_167 c5 c5small > gpurun_out/r2_bench_configs_c.jsonl 2> gpurun_out/r2_bench_configs_c.err; echo rc=$?
tail -c 800 gpurun_out/r2_bench_configs_c.err
python - <<PY
import json
for l in open("gpurun_out/r2_bench_configs_c.jsonl"):
    d=json.loads(l)
    if "enc_ms" in d: print(d["config"], d["mode"], d["rows"], "enc %.3f dec %.3f"%(d["enc_ms"],d["dec_ms"]), "%.1f M ct/s frac %.3f"%(d["ct_per_s"]/1e6,d["frac_hbm"]), d.get("matches_cuda_core_schedule"))
    else: print(d)
PY

#!/bin/bash
# kernel iteration: GPU tests (hang guard), then same-key timings at every BASELINE N
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_test_b.log; cat gpurun_out/r2_test_b.log
timeout 600 python scripts/bench_configs.py c1 c2 c3s c4 c5s > gpurun_out/r2_bench_configs_b.jsonl 2> gpurun_out/r2_bench_configs_b.err; echo rc=$?
tail -c 800 gpurun_out/r2_bench_configs_b.err
python - <<PY
import json
for l in open("gpurun_out/r2_bench_configs_b.jsonl"):
    d=json.loads(l); print(d["config"], d["rows"], "enc %.3f dec %.3f"%(d["enc_ms"],d["dec_ms"]), d["kernel_ms"], "%.1f M ct/s frac %.3f"%(d["ct_per_s"]/1e6,d["frac_hbm"]), d["roundtrip_equals_message"])
PY

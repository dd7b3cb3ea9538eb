#!/bin/bash
# first GPU validation of round 2: tests, smoke, bench with all configs
nvidia-smi -L; nproc; lscpu | grep -E "Model name" 
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_test_a.log; cat gpurun_out/r2_test_a.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo bench rc=$?
tail -c 1500 gpurun_out/r2_bench_a.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_a.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],d["e2e"].get("pcie_roof"),d["e2e"].get("frac_of_pcie"))
print(json.dumps(d["roofline"]["kernels"]))
print(json.dumps(d["extras"]["configs"],indent=1))
PY

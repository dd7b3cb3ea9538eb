#!/bin/bash
# N GPUs (first argument): bench.py under torchrun with every BASELINE config, as the driver launches it
N=${1:-8}
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo bench rc=$?
tail -c 300 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"].get("frac_of_pcie"),"plain",d["e2e"]["plain_arrays"]["value"])
for k,v in d["extras"]["configs"].items(): print(k, {x:v[x] for x in v if x in ("ms","ct_per_s","frac_hbm","matches_int64_column_sums","scaling")})
PY

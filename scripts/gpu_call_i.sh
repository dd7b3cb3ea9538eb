#!/bin/bash
# two GPUs: exchange stress, the world-2 hardware test, bench.py under torchrun with every config
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/xchg_stress.py 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "world2 or host_buffer_pack or device_drawn or sampler" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo bench rc=$?
tail -c 600 gpurun_out/r2_bench_2gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_2gpu.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],d["e2e"].get("pcie_roof"),d["e2e"].get("frac_of_pcie"))
for k,v in d["extras"]["configs"].items(): print(k, {x:v[x] for x in v if x not in ("workload","kernel_ms","checksum")})
PY

#!/bin/bash
# epilogue as one group (all warps on every phase): bit-equality with the two-group epilogue, then timing per kernel
nvidia-smi -L
timeout 300 python - <<'PY' 2>&1 | tail -8
import sys, numpy as np, torch
sys.path.insert(0, ".")
import ntru_circom_b200 as nb
for cfg in ("default167", "hps509", "hps677", "hps821", "hrss701"):
    g = dict(np.load(f"tests/golden/{cfg}.npz"))
    N, q = int(g["N"]), int(g["q"])
    eng = nb.Engine(N, 3, q, 0); eng.set_public_key(g["h"]); eng.set_private_key(g["f"], g["fp"])
    eng.set_stream(torch.cuda.current_stream().cuda_stream); eng.set_path(nb.PATH_TENSOR)
    B, P = 74 * 256 * 2 + 99, eng.pitch
    gen = torch.Generator(device="cuda").manual_seed(N)
    r = torch.zeros((B, P), dtype=torch.uint8, device="cuda"); r[:, :N] = torch.randint(0, 3, (B, N), generator=gen, device="cuda", dtype=torch.uint8)
    m = torch.zeros((B, P), dtype=torch.uint8, device="cuda"); m[:, :N] = torch.randint(0, 2, (B, N), generator=gen, device="cuda", dtype=torch.uint8)
    outs = []
    for og in (0, 15):
        eng.set_option(9, og)
        bufs = [torch.full((B, P), 7, dtype=torch.int16, device="cuda") for _ in range(4)] + [torch.full((B, P), 7, dtype=torch.uint8, device="cuda") for _ in range(2)]
        eng.encrypt_dev(B, r, m, value=bufs[0], quotientE=bufs[1])
        eng.decrypt_dev(B, bufs[0], value=bufs[4], quotient1=bufs[2], remainder1=bufs[3], quotient2=bufs[5])
        eng.sync(); outs.append(bufs)
    print(cfg, "one-group == two-group:", all(torch.equal(x, y) for x, y in zip(*outs)))
    eng.close()
PY
for OG in 0 15 0 15; do
ONEGROUP=$OG timeout 600 python scripts/bench_configs.py c1 c2 c3s c4 c5s 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('ONEGROUP=$OG', d['config'], d['kernel_ms'], '%.1f M ct/s' % (d['ct_per_s'] / 1e6))
" | tee -a gpurun_out/r2_one_group.txt
done

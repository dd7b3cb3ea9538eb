"""Small tcgen05-schedule run for compute-sanitizer (one tool per gpurun call):
  compute-sanitizer --tool racecheck|initcheck|memcheck|synccheck python scripts/sanitize_target.py [N q rows]
Several 256-row tiles per CTA pair are not needed for the hazards these tools see (shared-memory races, reads of
uninitialised global memory, out-of-bounds accesses), so the batch is two waves of tiles plus a ragged tail; outputs are
poisoned first and compared with the fp32 schedule afterwards, pad columns included (the round-1 report of a `value`
array left unwritten at N = 701 is what this chases)."""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import ntru_circom_b200 as nb  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 701
q = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2 * 256 + 131
p, dev = 3, "cuda"
rng = np.random.default_rng(N)
eng = nb.Engine(N, p, q, 0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.set_public_key(rng.integers(0, q, size=N).astype(np.uint16))
eng.set_private_key(rng.integers(-1, 2, size=N).astype(np.int8), rng.integers(0, p, size=N).astype(np.uint8))
P = eng.pitch
r = torch.zeros((B, P), dtype=torch.uint8, device=dev)
r[:, :N] = torch.randint(0, 3, (B, N), device=dev, dtype=torch.uint8)
m = torch.zeros((B, P), dtype=torch.uint8, device=dev)
m[:, :N] = torch.randint(0, 2, (B, N), device=dev, dtype=torch.uint8)
res = {}
for path in (nb.PATH_TENSOR, nb.PATH_CUDA_CORE):
    eng.set_path(path)
    bufs = [torch.full((B, P), 7, dtype=torch.int16, device=dev) for _ in range(4)] + \
           [torch.full((B, P), 7, dtype=torch.uint8, device=dev) for _ in range(2)]
    val, quo, q1, r1, pv, q2 = bufs
    for _ in range(2 if path == nb.PATH_TENSOR else 1):
        eng.encrypt_dev(B, r, m, value=val, quotientE=quo)
        eng.decrypt_dev(B, val, value=pv, quotient1=q1, remainder1=r1, quotient2=q2)
    eng.sync()
    res[path] = bufs
ok = all(torch.equal(a, b) for a, b in zip(res[nb.PATH_TENSOR], res[nb.PATH_CUDA_CORE]))
print("sanitize_target", N, q, B, "tensor == fp32 schedule (pad columns included):", ok)
eng.close()
sys.exit(0 if ok else 1)

"""Roofline evidence from an ncu report, written by script (no hand transcription):
  python scripts/ncu_traffic.py <rep.ncu-rep> <rows per launch> <N> <q> <out.json> [label]
Per captured k_umma_pair launch: measured DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per ciphertext next
to the algorithmic bytes, the wasted-traffic ratio, tensor-pipe activity, and executed / algorithmic int8
MMA work (executed = what the chunk table and the K-atom skipping of umma_kernels.cu issue, recomputed here from N, q).
The report comes from `ncu --set full --clock-control none -k regex:k_umma_pair` around scripts/profile_target.py."""
import csv
import io
import json
import subprocess
import sys

rep, rows, N, q, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
label = sys.argv[6] if len(sys.argv) > 6 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
table = list(csv.reader(io.StringIO(raw)))
hdr, data = table[0], table[2:]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return None


def geometry(mode):
    """chunk table of umma_kernels.cu:geometry()"""
    nl = 2 if (mode == "enc" and q > 256) else 1
    kp = (N + 127) // 128 * 128
    max_out = 128 if mode == "enc" else 256 // nl
    pu1 = mode == "enc" and kp // 128 >= 5
    g = (32 if pu1 else 64) if mode == "enc" else (32 if mode == "dec1" else 128)
    T = (N + g - 1) // g * g
    n = (T + max_out - 1) // max_out
    base = T // n // g * g
    wide = (T - base * n) // g
    c0 = [0]
    for c in range(n):
        c0.append(c0[-1] + base + (g if c < wide else 0))
    return nl, kp, c0


def executed_macs(mode):
    """int8 MACs per ciphertext the kernel issues: the phase list of umma_kernels.cu:build_schedule() (hi product, then the
    lo product on top of it, chunks in pairs; an odd last chunk as cyclic + hi), columns x K bytes actually multiplied"""
    nl, kp, c0 = geometry(mode)
    kl = 2 if (mode == "dec1" and q > 256) else 1
    atoms = kp // 128
    k_last = (N - (atoms - 1) * 128 + 31) // 32 * 32
    n = len(c0) - 1

    def kbytes(a0, a1):           # atoms [a0, a1): the last atom of the operand holds k_last bytes
        return sum(k_last if at == atoms - 1 else 128 for at in range(a0, a1))

    def hi_a0(c):
        return min((c0[c] + 1) // 128, atoms - 1)

    def lo_a1(c):
        return (min(c0[c + 1], N) - 1) // 128 + 1

    phases = []
    c = n - 1
    if n & 1:
        phases += [(c, 0, atoms), (c, hi_a0(c), atoms)]
        c -= 1
    while c >= 1:
        phases += [(c, hi_a0(c), atoms), (c - 1, hi_a0(c - 1), atoms), (c, 0, lo_a1(c)), (c - 1, 0, lo_a1(c - 1))]
        c -= 2
    return sum(nl * (c0[c + 1] - c0[c]) * kbytes(a0, a1) * kl for c, a0, a1 in phases)


limbs = 2 if q > 256 else 1
alg_bytes = {"enc": 6 * N, "dec1": 6 * N, "dec2": 2 * N}
alg_macs = {"enc": N * N * limbs, "dec1": N * N * limbs, "dec2": N * N}
modes = {"ILi0E": "enc", "ILi1E": "dec1", "ILi2E": "dec2"}
res = {"source": f"ncu --set full --clock-control none, {rep}, {rows} rows per launch (inputs + outputs exceed the 126 MB L2)", "label": label,
       "N": N, "q": q, "rows": rows, "kernels": {}, "bytes_per_ciphertext": {}}
for r in data:
    name = r[col["Kernel Name"]] if "Kernel Name" in col else ""
    if "k_umma_pair" not in name:
        continue
    mode = None
    for tag, m in (("k_umma_pair<0", "enc"), ("k_umma_pair<1", "dec1"), ("k_umma_pair<2", "dec2"), ("(Mode)0", "enc"), ("(Mode)1", "dec1"), ("(Mode)2", "dec2")):
        if tag in name.replace(" ", ""):
            mode = m
            break
    if mode is None:
        for tag, m in modes.items():
            if tag in name:
                mode = m
    if mode is None or mode in res["kernels"]:
        continue
    rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
    unit_rd, unit_wr = table[1][col["dram__bytes_read.sum"]], table[1][col["dram__bytes_write.sum"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd, wr = rd * scale.get(unit_rd, 1), wr * scale.get(unit_wr, 1)
    ex = executed_macs(mode)
    tscale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}
    k = {"duration_us_under_ncu": num(r, "gpu__time_duration.sum") * tscale.get(table[1][col["gpu__time_duration.sum"]], 1.0),
         "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_ciphertext": (rd + wr) / rows,
         "algorithmic_bytes_per_ciphertext": alg_bytes[mode], "traffic_over_algorithmic": (rd + wr) / rows / alg_bytes[mode],
         "dram_throughput_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
         "tensor_pipe_active_pct": num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
         "lts_throughput_pct": num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
         "registers_per_thread": num(r, "launch__registers_per_thread"),
         "executed_int8_macs_per_ciphertext": ex, "algorithmic_int8_macs_per_ciphertext": alg_macs[mode],
         "executed_over_algorithmic_mma": ex / alg_macs[mode]}
    res["kernels"][mode + "_tensor"] = k
    res["bytes_per_ciphertext"][mode + "_tensor"] = (rd + wr) / rows
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))

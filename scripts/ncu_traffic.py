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


def fp16_dec1():
    """the first decrypt product runs on kind::f16 tiles (umma_prepare_private: 256 < q <= 2048, default form)"""
    return 256 < q <= 2048


def one_group(mode):
    """umma_kernels.cu:one_group_mode(), default NTRU_OPT_EPILOGUE"""
    if mode == "dec1":
        return N <= 512 if fp16_dec1() else N <= 768
    if mode == "dec2":
        return N <= 512 and (N + 255) // 256 * 256 == (N + 127) // 128 * 128
    return False


def geometry(mode):
    """chunk table of umma_kernels.cu:geometry(); ea = coefficients per 128-byte K atom"""
    f16 = mode == "dec1" and fp16_dec1()
    nl = 2 if (mode == "enc" and q > 256) else 1
    ea = 64 if f16 else 128
    atoms = (N + ea - 1) // ea
    max_out = 128 if mode == "enc" else 256 // nl
    pu1 = mode == "enc" and atoms >= 5
    g = (32 if pu1 else 64) if mode == "enc" else ((64 if f16 else 32) if mode == "dec1" else 128)
    if one_group(mode):
        g *= 2
    T = (N + g - 1) // g * g
    n = (T + max_out - 1) // max_out
    base = T // n // g * g
    wide = (T - base * n) // g
    c0 = [0]
    for c in range(n):
        c0.append(c0[-1] + base + (g if c < wide else 0))
    return nl, ea, atoms, c0


def executed_macs(mode):
    """int8 MACs (int8-equivalents for the fp16 form: one fp16 MAC takes the tensor time of two int8 MACs) per ciphertext
    the kernel issues: the phase list of umma_kernels.cu:build_schedule() (hi product, then the lo product on top of it,
    chunks in pairs; an odd last chunk as cyclic + hi), columns x K coefficients actually multiplied, in 32-byte steps"""
    nl, ea, atoms, c0 = geometry(mode)
    f16 = mode == "dec1" and fp16_dec1()
    kl = 2 if (mode == "dec1" and q > 256 and not f16) else 1
    step = ea // 4
    k_last = (N - (atoms - 1) * ea + step - 1) // step * step
    n = len(c0) - 1

    def kcoef(a0, a1):           # atoms [a0, a1): the last atom of the operand holds k_last coefficients
        return sum(k_last if at == atoms - 1 else ea for at in range(a0, a1))

    def hi_a0(c):
        return min((c0[c] + 1) // ea, atoms - 1)

    def lo_a1(c):
        return (min(c0[c + 1], N) - 1) // ea + 1

    phases = []
    c = n - 1
    if n & 1:
        phases += [(c, 0, atoms), (c, hi_a0(c), atoms)]
        c -= 1
    while c >= 1:
        phases += [(c, hi_a0(c), atoms), (c - 1, hi_a0(c - 1), atoms), (c, 0, lo_a1(c)), (c - 1, 0, lo_a1(c - 1))]
        c -= 2
    return sum(nl * (c0[c + 1] - c0[c]) * kcoef(a0, a1) * kl for c, a0, a1 in phases) * (2 if f16 else 1)


limbs = 2 if q > 256 else 1
alg_bytes = {"enc": 6 * N, "dec1": 6 * N, "dec2": 2 * N}
alg_macs = {"enc": N * N * limbs, "dec1": N * N * limbs, "dec2": N * N}
modes = {"ILi0E": "enc", "ILi1E": "dec1", "ILi2E": "dec2", "ILi3E": "dec1"}
res = {"source": f"ncu --set full --clock-control none, {rep}, {rows} rows per launch (inputs + outputs exceed the 126 MB L2)", "label": label,
       "N": N, "q": q, "rows": rows, "kernels": {}, "bytes_per_ciphertext": {}}
for r in data:
    name = r[col["Kernel Name"]] if "Kernel Name" in col else ""
    if "k_umma_pair" not in name:
        continue
    mode = None
    for tag, m in (("k_umma_pair<0", "enc"), ("k_umma_pair<1", "dec1"), ("k_umma_pair<2", "dec2"), ("k_umma_pair<3", "dec1"),
                   ("k_umma_pair<(int)0", "enc"), ("k_umma_pair<(int)1", "dec1"), ("k_umma_pair<(int)2", "dec2"), ("k_umma_pair<(int)3", "dec1"),
                   ("(Mode)0", "enc"), ("(Mode)1", "dec1"), ("(Mode)2", "dec2"), ("(Mode)3", "dec1")):
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

"""Summarise an .ncu-rep: python scripts/ncu_summary.py rep [--stalls KERNEL_INDEX]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_uniform.sum", "lts__t_sectors_srcunit_tex.sum"]
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{units[i]}]:", [r[i][:60] for r in rows[2:]])
if "--stalls" in sys.argv:
    kidx = int(sys.argv[sys.argv.index("--stalls") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    kern, cur = [], None
    for r in srows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    k = kern[kidx]
    h, data = k["rows"][0], k["rows"][1:]
    isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    sc = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[isamp] or 0) for r in data)
    print(k["name"], "samples", tot)
    top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:int(sys.argv[-1]) if sys.argv[-1].isdigit() else 30]
    for i in sorted(top):
        r = data[i]
        st = {h[c][6:]: int(r[c]) for c in sc if r[c] and int(r[c]) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(i, r[isamp], r[iex], r[isrc][:80], st)

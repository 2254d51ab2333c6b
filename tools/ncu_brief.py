"""Short markdown summary of an .ncu-rep (read on the CPU box with `ncu -i`): headline raw metrics per captured
launch + executed-instruction mix and the hottest source lines.  python tools/ncu_brief.py rep out.md [units_per_launch name]"""
import collections
import csv
import io
import subprocess
import sys

rep, out_path = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else None
uname = sys.argv[4] if len(sys.argv) > 4 else "unit"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit_row = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = [f"# {rep.split('/')[-1]}", "", "`ncu --set full --clock-control none --import-source on` on a B200; per-launch values (cold-ish caches, serialised).", ""]
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, unit_row, vals)}
    out.append(f"## {d.get('Kernel Name', ('?', ''))[0][:110]}")
    for k in KEYS:
        if k in d and d[k][0] != "":
            out.append(f"- {k} = {d[k][0]} {d[k][1]}")
    if units and "smsp__inst_executed.sum" in d:
        out.append(f"- warp instructions per {uname} = {float(d['smsp__inst_executed.sum'][0]) / units:.1f}")
    out.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
try:
    hi = next(i for i, r in enumerate(rows[:6]) if "Source" in r)
    h = rows[hi]
    iS, iE, iSm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    data, prev = [], -1
    for r in rows[hi + 1:]:
        if len(r) <= iE:
            continue
        try:
            a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
        except ValueError:
            continue
        if a < prev:
            break
        prev = a
        data.append(r)
    tot = sum(int(r[iE] or 0) for r in data)
    hist = collections.Counter()
    for r in data:
        t = r[iS].strip().split()
        if t:
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            hist[op.split(".")[0]] += int(r[iE] or 0)
    out.append(f"### executed warp instructions of the first captured launch: {tot}")
    out.append(", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in hist.most_common(14)))
    sm = sum(int(r[iSm] or 0) for r in data)
    out.append(f"\n### warp-stall samples (total {sm}); SASS lines with >= 1.5% of the samples")
    for idx, r in enumerate(data):
        s = int(r[iSm] or 0)
        if sm and s >= 0.015 * sm:
            out.append(f"- [{idx}] {100 * s / sm:.1f}%  {r[iS].strip()[:110]}")
except Exception as e:  # noqa: BLE001
    out.append(f"(no source page: {e})")
open(out_path, "w").write("\n".join(out) + "\n")
print("\n".join(out[:60]))

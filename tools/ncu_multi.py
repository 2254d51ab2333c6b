"""Summarise an .ncu-rep that holds SEVERAL kernel launches (read here on the CPU box with `ncu -i`): per launch the key
raw metrics, then - per distinct kernel - where the warp-stall samples and the executed instructions sit in the SASS
(100-instruction regions + the hottest single instructions).  Usage: python tools/ncu_multi.py x.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = [f"# {rep.split('/')[-1]}: {len(rows) - 2} launches (ncu --set full --clock-control none)", ""]
names = []
for n, vals in enumerate(rows[2:]):
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    name = d.get("Kernel Name", ("?", ""))[0]
    names.append(name)
    out.append(f"## launch {n}: {name[:100]}")
    for k in KEYS:
        if k in d:
            out.append(f"- {k} = {d[k][0]} {d[k][1]}")
seen = set()
for n, name in enumerate(names):
    if name in seen:
        continue
    seen.add(name)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-s", str(n), "-c", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if not starts:
        continue
    s = starts[0]
    e = starts[1] if len(starts) > 1 else len(rows)
    h = rows[s + 1]
    data = [r for r in rows[s + 2:e] if len(r) == len(h)]
    iS, iSrc, iE = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    tot = max(1, sum(int(r[iS]) for r in data))
    tote = max(1, sum(int(r[iE]) for r in data))
    out.append(f"\n### launch {n} ({name[:60]}): {len(data)} SASS instructions, {tote} warp instructions executed, {tot} stall samples")
    out.append("| SASS index | % samples | % executed | first instruction |")
    out.append("|---|---|---|---|")
    for i in range(0, len(data), 100):
        ss = sum(int(r[iS]) for r in data[i:i + 100])
        ee = sum(int(r[iE]) for r in data[i:i + 100])
        if ss or ee:
            out.append(f"| {i}-{i + 99} | {100 * ss / tot:.1f} | {100 * ee / tote:.1f} | `{data[i][iSrc].strip()[:60]}` |")
    out.append("\nhottest instructions:")
    for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][iS]))[:12]:
        out.append(f"- [{idx}] {100 * int(r[iS]) / tot:.1f}%  `{r[iSrc].strip()[:90]}`")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")

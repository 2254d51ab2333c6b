"""Summarise an .ncu-rep (read here on the CPU box with `ncu -i`): key raw metrics + stall samples
per warp-role region of the fused kernel.  Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_op_read_hit_rate.pct",
        "lts__t_sector_op_write_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sectors_srcunit_ltcfabric.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out.append(f"## {d.get('Kernel Name', ('?', ''))[0][:90]}")
    for k in KEYS:
        if k in d:
            out.append(f"- {k} = {d[k][0]} {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
try:
    h = rows[1]; data = rows[2:]
    iS, iSrc = h.index("# Samples"), h.index("Source")
    tot = sum(int(r[iS]) for r in data)
    out.append(f"\n### warp-stall samples (total {tot}); instructions with >= 0.8% of samples")
    for idx, r in enumerate(data):
        s = int(r[iS])
        if s >= 0.008 * tot:
            out.append(f"- [{idx}] {100 * s / tot:.1f}%  {r[iSrc].strip()[:100]}")
except Exception as e:  # noqa: BLE001
    out.append(f"(no source page: {e})")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")

"""DRAM traffic of ONE launch of the fused inference kernel as timed by bench.py (2^20 poses), from an ncu metrics pass:
    python tools/capture_traffic.py              (on the GPU box; writes gpurun_out/mlp_tc_traffic.json)
bench.py reports it as roofline.traffic and refuses a file captured from another version of csrc/mlp_tc.cu."""
import csv, datetime, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "gpurun_out"); os.makedirs(out_dir, exist_ok=True)
log = os.path.join(out_dir, "mlp_tc_traffic.csv")
cmd = ["ncu", "--metrics", "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
       "-k", "regex:mlp_forward_tc", "-s", "1", "-c", "3", "--csv", "--log-file", log, sys.executable, os.path.join(ROOT, "tools", "forward_once.py"), str(1 << 20), "3"]
subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
rd, wr, ns = [], [], []
for r in csv.reader(open(log)):
    if len(r) > 10 and r[0].isdigit():
        v = float(r[-1].replace(",", ""))
        {"dram__bytes_read.sum": rd, "dram__bytes_write.sum": wr, "gpu__time_duration.sum": ns}[r[-3]].append(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(r[-2], 1))
with open(os.path.join(ROOT, "3d-pose-baseline_b200", "csrc", "mlp_tc.cu"), "rb") as f:
    sha = hashlib.sha256(f.read()).hexdigest()[:16]
res = {"kernel": "mlp_forward_tc_kernel<2>", "kernel_src_sha16": sha, "when": datetime.datetime.utcnow().strftime("%Y-%m-%dT%H:%MZ"),
       "capture": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, launches 2-4 of tools/forward_once.py 1048576 (mean)",
       "launches": len(rd), "dram_bytes_read": sum(rd) / len(rd), "dram_bytes_write": sum(wr) / len(wr),
       "dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd), "kernel_ns_under_ncu": sum(ns) / len(ns),
       "algorithmic_bytes_per_launch": (1 << 20) * 320, "l2_persist": os.environ.get("P3D_L2_PERSIST", "1")}
json.dump(res, open(os.path.join(out_dir, "mlp_tc_traffic.json"), "w"), indent=1)
print(json.dumps(res))

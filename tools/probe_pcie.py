"""PCIe probe: pinned H2D / D2H bandwidth of GPU 0 with the allocating thread bound to each NUMA node, alone and both
directions at once (what the end-to-end path of bench.py does).  Usage: python tools/probe_pcie.py"""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def nodes():
    out = {}
    base = "/sys/devices/system/node"
    if not os.path.isdir(base):
        return out
    for d in sorted(os.listdir(base)):
        if d.startswith("node") and d[4:].isdigit():
            cpus = open(os.path.join(base, d, "cpulist")).read().strip()
            s = set()
            for part in cpus.split(","):
                if "-" in part:
                    a, b = part.split("-"); s.update(range(int(a), int(b) + 1))
                elif part:
                    s.add(int(part))
            out[int(d[4:])] = s
    return out


def bw(nbytes=256 << 20, iters=6):
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for name, do_in, do_out in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            if do_in:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        res[name] = round(nbytes * iters / (ms * 1e-3) / 1e9, 1)        # GB/s per direction
    return res


if __name__ == "__main__":
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500], file=sys.stderr)
    except Exception:
        pass
    torch.cuda.init()
    nd = nodes()
    allc = os.sched_getaffinity(0)
    out = {"cpus_allowed": len(allc), "numa_nodes": {k: len(v) for k, v in nd.items()}, "default": bw()}
    for k, cpus in nd.items():
        use = cpus & allc
        if not use:
            continue
        os.sched_setaffinity(0, use)
        out["node%d" % k] = bw()
    os.sched_setaffinity(0, allc)
    print(json.dumps(out))

"""The exchanges of a data-parallel training step on their own (launch with torchrun): the flat fp32 gradient all-reduce
(NCCL) and one SyncBN-sized sum over peer memory, CUDA events, max over ranks; plus the whole step at three global
batch sizes.  Run under different NCCL_* settings to see what the gradient exchange can be made to cost."""
import ctypes as C, json, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-pose-baseline_b200")]
from p3d import LinearModel, _lib
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def timed(fn, iters, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) * 1e3
sptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = {"world": world, "nccl_env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}
for Bg in [int(b) for b in (sys.argv[1:] or [4096])]:
    m = LinearModel(1024, 2, True, True, True, Bg, 1e-3, mode="bf16", device=local, seed=1, dist=dist)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((Bg, 32), device=dev, generator=g); t = torch.randn((Bg, 48), device=dev, generator=g)
    out["step_us_B%d" % Bg] = round(timed(lambda: m.step(None, x, t, 0.5, isTraining=True), 20), 1)
    out["grad_allreduce_us"] = round(timed(lambda: _lib.check(_lib.lib.p3d_debug_dp_part(m._handle, 0, sptr)), 30), 1)
    out["one_exchange_us"] = round(timed(lambda: _lib.check(_lib.lib.p3d_debug_dp_part(m._handle, 1, sptr)), 50), 1)
    m.close()
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()

"""Host <-> device copy bandwidth of all ranks at once (torchrun, one rank per GPU, NUMA-pinned like bench.py): the ceiling of the
`e2e` number at N GPUs.  usage: python -m torch.distributed.run --nproc-per-node N tools/probe/pcie_probe.py"""
import os, time, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
try:
    import pynvml
    pynvml.nvmlInit(); pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
except Exception as e:
    print("affinity unavailable", e)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
h_out = torch.empty(n // 2, dtype=torch.uint8).pin_memory(); h_out.fill_(0)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.ones(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=4):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    dt = time.perf_counter() - t0
    return (n * reps * h2d) / dt / 1e9, (n // 2 * reps * d2h) / dt / 1e9
run(True, True, 1)
for name, a, b in (("h2d only", True, False), ("d2h only", False, True), ("both", True, True)):
    x, y = run(a, b)
    t = torch.tensor([x, y], device="cuda")
    if world > 1: dist.all_reduce(t)
    if local == 0: print(f"{name}: per rank h2d {x:.1f} d2h {y:.1f} GB/s; all ranks h2d {t[0].item():.1f} d2h {t[1].item():.1f} GB/s", flush=True)
if local == 0:
    os.system("nvidia-smi topo -m | head -14; lscpu | grep -E 'Model name|Socket|NUMA node|^CPU\\(s\\)' | head -8; free -g | head -2")

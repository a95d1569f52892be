import os, time, torch, subprocess
print(subprocess.run("nvidia-smi topo -m | head -8; lscpu | grep -i 'numa\|^CPU(s)'", shell=True, capture_output=True, text=True).stdout)
import pynvml as nv
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
n = (os.cpu_count() + 63) // 64
mask = nv.nvmlDeviceGetCpuAffinity(h, n)
cpus = [i * 64 + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
print("gpu0 ideal cpus:", cpus[:8], "...", len(cpus), "current affinity:", len(os.sched_getaffinity(0)))
def bw(tag):
    a = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2): d.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10): d.copy_(a, non_blocking=True)
    torch.cuda.synchronize(); h2d = 2560 / (time.perf_counter() - t) / 1024
    t = time.perf_counter()
    for _ in range(10): a.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); d2h = 2560 / (time.perf_counter() - t) / 1024
    print(f"{tag}: H2D {h2d:.1f} GiB/s  D2H {d2h:.1f} GiB/s")
bw("default placement")
allc = os.sched_getaffinity(0)
far = sorted(set(allc) - set(cpus))
if far:
    os.sched_setaffinity(0, far); bw("far cpus")
os.sched_setaffinity(0, [c for c in cpus if c in allc] or allc); bw("gpu-local cpus")

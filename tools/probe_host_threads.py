"""Developer probe: how many host threads make the CPU reference arm fastest on this box?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(p, open(p).read().strip())
    except Exception as e: print(p, "n/a")
try: print("loadavg", open("/proc/loadavg").read().strip())
except Exception: pass
import torch
import bench
for th in (4, 8, 16, 32, 64, 128):
    t0 = time.time()
    v, sec, samples, cores = bench.cpu_reference_samples_per_s(1, steps=1, warmup=0, threads=th)
    print(f"threads={th:4d}: {v:10.0f} samples/s  ({samples} samples in {sec:.2f} s; wall {time.time()-t0:.1f} s)", flush=True)

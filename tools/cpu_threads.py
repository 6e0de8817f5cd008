"""Pick the host thread count that makes the CPU port fastest (bench.py's cpu_baseline)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
for th in (8, 16, 32, 64, 128):
    if th > (os.cpu_count() or 1):
        continue
    step = bench.cpu_port_rate(1024, th)
    step()
    t0 = time.perf_counter(); step(); dt = time.perf_counter() - t0
    print(f"threads={th}: {1024 / dt:.1f} rays/s", flush=True)

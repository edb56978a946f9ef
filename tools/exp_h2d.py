"""Aggregate pinned host-to-device bandwidth with one process per GPU (no kernels): python tools/exp_h2d.py NGPUS"""
import subprocess, sys, time
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
child = r'''
import torch, time, sys
dev = int(sys.argv[1]); torch.cuda.set_device(dev)
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0 = float(sys.argv[2])
while time.time() < t0: pass
t = time.perf_counter()
for _ in range(40): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
print(dev, round(40 * (256 << 20) / (time.perf_counter() - t) / 1e9, 1), flush=True)
'''
t0 = time.time() + 25
ps = [subprocess.Popen([sys.executable, "-c", child, str(i), str(t0)], stdout=subprocess.PIPE, text=True) for i in range(n)]
vals = [float(p.communicate()[0].split()[1]) for p in ps]
print(f"{n} GPUs: per-GPU GB/s {vals} aggregate {sum(vals):.1f}")

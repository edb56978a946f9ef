python - <<'PY'
import torch, time
x = torch.empty(256<<20, dtype=torch.uint8).pin_memory()
d = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(10): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt=time.perf_counter()-t
print('pinned H2D GB/s', 10*(256<<20)/dt/1e9)
PY
for p in 32 48 64; do CUDA_DEVICE_MAX_CONNECTIONS=32 python bench.py --no-cpu-baseline --steps 4 --warmup 2 --batch 256 --pipelines $p | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pipelines', d['config']['pipelines_per_rank'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2))"; done

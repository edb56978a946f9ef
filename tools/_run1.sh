for cs in 1 0; do JXLB200_COPY_STREAM=$cs python bench.py --no-cpu-baseline --steps 5 --warmup 3 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('copy_stream $cs value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'GB/s', round(d['e2e']['h2d_bytes_per_step']/d['e2e']['ms_per_step']/1e6,1))"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch or strided or error" 2>&1 | tail -2

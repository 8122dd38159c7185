#!/bin/bash
# e2e (host-buffer) arm of the bench for several host-pipeline chunk sizes
for mb in 8 16 32 64 128; do
  CB_HOST_CHUNK_MB=$mb python bench.py --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk_mb', $mb, 'e2e GB/s', round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],2))"
done

#!/bin/bash
# A/B of density variants on C5 (development): usage scripts/c5_ab.sh [variant ...]
one() { python bench.py --configs c5 --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['configs']['C5']; print('$1', round(c['ms'],3), round(c['ms_min'],3), c['parity']['max_rel'])"; }
one default
for v in "$@"; do CB_CUDA_LIB=$PWD/variants/lib$v.so one $v; done

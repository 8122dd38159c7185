#!/bin/bash
# A/B of banded-apply pipeline variants (development): usage scripts/k7_ab.sh [variant ...]
one() { python bench.py --configs kernels 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('$1', 'c1', round(k['c1_banded_apply_65536vec']['us'],1), k['c1_banded_apply_65536vec']['parity_max_rel'], 'c3', round(k['c3_banded_apply_256vec']['us'],1), k['c3_banded_apply_256vec']['parity_max_rel'])"; }
one default
for v in "$@"; do CB_CUDA_LIB=$PWD/variants/lib$v.so one $v; done

#!/bin/bash
# A/B of the FE-DVR assembly variants on the headline step (development)
one() { python bench.py --no-extras --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['ms_per_step']*1e3,2), round(d['roofline']['kernel_ms']*1e3,2), d['config']['k4_element_shards']['us_full'])"; }
one default
CB_FEDVR_NO_PDL=1 one no_pdl
for v in "$@"; do CB_CUDA_LIB=$PWD/variants/lib$v.so one $v; done
one default_again

#!/bin/bash
# A/B matrix of env-kernel variants with benchmarks/tail_probe.py (50 000 and 4 000 rows)
run() { echo "== $*"; env "$@" python benchmarks/tail_probe.py --rows 50000 4000 2>&1 | tail -2; }
run TTL_K1_MODE=quad
run TTL_K1_MODE=row TTL_K1_MINB=4
run TTL_K1_MODE=row TTL_K1_MINB=6
run TTL_K1_MODE=row TTL_K1_MINB=8
run TTL_K1_MODE=row TTL_K1_MINB=4 TTL_STATE_MINB=4
run TTL_K1_MODE=row TTL_K1_MINB=4 TTL_STATE_MINB=8

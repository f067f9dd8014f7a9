#!/bin/bash
# headline circuit only (no CPU / e2e / strong-scaling / other-config legs): prints ms_per_step for each QCS_TC_DEBUG value given
for d in "$@"; do
  QCS_TC_DEBUG=$d python bench.py --steps 3 --warmup 2 --skip-cpu --skip-e2e --skip-strong --skip-configs 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dbg $d', l['ms_per_step'], l['value'])"
done

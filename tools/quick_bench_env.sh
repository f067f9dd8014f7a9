#!/bin/bash
# headline circuit only, once per environment setting given as arguments ("VAR=value" or "-" for the defaults)
for kv in "$@"; do
  if [ "$kv" = "-" ]; then e=""; else e="$kv"; fi
  env $e python bench.py --steps 3 --warmup 2 --skip-cpu --skip-e2e --skip-strong --skip-configs 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$kv', round(l['ms_per_step'],2), round(l['value']), l['roofline']['plan'])"
done

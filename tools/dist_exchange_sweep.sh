#!/bin/bash
# exchange kernel tuning: unroll x CTAs per SM (one process, all GPUs)
for u in 8 16; do for c in 4 8 16; do
  echo "unroll $u ctas $c: $(QCS_DIST_UNROLL=$u QCS_DIST_CTAS=$c python tools/dist_exchange_check.py single 2>&1 | grep exchange)"
done; done

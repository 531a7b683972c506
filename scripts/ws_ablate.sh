#!/bin/bash
# forward-only bench of the warp-specialised kernel with roles switched off (GNN_B200_WS_DEBUG: 1 no row copies, 2 no segment
# sums, 4 no mma): which role bounds the iteration?  Results are WRONG with any bit set; timing only.
for d in 0 1 2 4 3 6 7; do
  echo "== WS_DEBUG=$d"
  GNN_B200_WS_DEBUG=$d timeout 200 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-train --skip-variant --skip-parity 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        r = json.loads(line); print('   ms_per_launch', r['roofline']['ms_per_launch'], 'frac', r['roofline']['frac'], 'step ms', r['ms_per_step'])
"
done

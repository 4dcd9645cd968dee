#!/bin/bash
# A/B of tuning builds on one GPU box: usage tools/ab.sh tag lib1 [lib2 ...]  (libs relative to the repo root)
# prints ms per frame / march ms / samples for the bench frame and the two sample-bound stress frames
tag=$1; shift
for lib in "$@"; do
  for cfg in "" "--zoom 4" "--zoom 4 --regime translucent"; do
    NMR_LIB=$PWD/$lib python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline $cfg 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('$tag | $lib | $cfg | ms/frame %.4f (median %.4f, max %.4f) | march ms %.4f | Msamples/s (march) %.0f | e2e fps %.0f | e2e_u8 fps %.0f' % (d['ms_per_step'], d['ms_per_step_median'], d['ms_per_step_max'], r['kernel_ms_per_launch'], r['samples_per_launch'] / r['kernel_ms_per_launch'] / 1e3, d['e2e']['fps'], d['e2e_u8']['fps']))
"
  done
done

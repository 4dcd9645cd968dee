#!/bin/bash
# one-line A/B of the bench frame only: tools/ab1.sh tag lib [bench args]
tag=$1; lib=$2; shift; shift
NMR_LIB=$PWD/$lib python bench.py --steps 60 --warmup 5 --no-extras --no-cpu-baseline "$@" 2>/tmp/ab1.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('$tag | ms/frame %.4f (median %.4f, max %.4f) | march ms %.4f | e2e fps %.0f | e2e_u8 fps %.0f' % (d['ms_per_step'], d['ms_per_step_median'], d['ms_per_step_max'], r['kernel_ms_per_launch'], d['e2e']['fps'], d['e2e_u8']['fps']))
"

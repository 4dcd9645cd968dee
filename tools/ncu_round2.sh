#!/bin/bash
# ncu evidence of round 2 (one GPU): plain runs first, then the launch list of the bench command and `--set full` captures of the
# march kernel (bench frame and the sample-bound frame) and of the set-up kernel.  Reports land in gpurun_out/, summaries are made
# from them with tools/ncu_summary.py and committed under profiles/.
set -x
O=gpurun_out
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/r2_plain_bench.json 2> $O/r2_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/r2_ncu_launch.log 2>&1
export NMR_NO_OVERLAP=1     # the serial variant: CUDA events and ncu see the march kernel alone
python tools/profile_frame.py --frames 4 > $O/r2_pf_opaque.log 2>&1 || exit 1
python tools/profile_frame.py --frames 4 --zoom 4 --regime translucent > $O/r2_pf_translucent.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:march_kernel --launch-skip 2 -c 1 -f -o $O/r2_march_opaque python tools/profile_frame.py --frames 4 > $O/r2_ncu_o.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:march_kernel --launch-skip 2 -c 1 -f -o $O/r2_march_translucent python tools/profile_frame.py --frames 4 --zoom 4 --regime translucent > $O/r2_ncu_t.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:init_rays_kernel --launch-skip 2 -c 1 -f -o $O/r2_init_rays python tools/profile_frame.py --frames 4 > $O/r2_ncu_i.log 2>&1
python tools/profile_frame.py --frames 4 --zoom 4 --regime translucent --log2T 24 > $O/r2_pf_translucent_T24.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:march_kernel --launch-skip 2 -c 1 -f -o $O/r2_march_translucent_T24 python tools/profile_frame.py --frames 4 --zoom 4 --regime translucent --log2T 24 > $O/r2_ncu_t24.log 2>&1
for f in $O/r2_pf_opaque.log $O/r2_pf_translucent.log $O/r2_pf_translucent_T24.log; do tail -n 2 $f; done

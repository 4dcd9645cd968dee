# The measurement run behind profiles/ (one B200 box, ~3 minutes): GPU tests, bench (both arms), smoke, ncu launch lists, three
# `ncu --set full` captures, the other BASELINE configs.  Outputs land in gpurun_out/r4_*; tools/ncu_summary.py turns the reports
# into profiles/r1_*_summary.json.   gpurun --timeout 2400 -- bash tools/final_measurements.sh
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r4_pytest.log 2>&1; tail -3 gpurun_out/r4_pytest.log
timeout 600 python bench.py > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err; tail -c 600 gpurun_out/r4_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r4_bench_ref.json 2> gpurun_out/r4_bench_ref.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4_smoke.log 2>&1; tail -2 gpurun_out/r4_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r4_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r4_launches_frame.csv python tools/profile_frame.py --frames 4 > gpurun_out/r4_ncu_frame.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:march_kernel<1, 0>" --launch-skip 3 -c 1 -f -o gpurun_out/r4_march_opaque python tools/profile_frame.py --frames 4 > gpurun_out/r4_ncu_o.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:march_kernel<1, 0>" --launch-skip 3 -c 1 -f -o gpurun_out/r4_march_translucent python tools/profile_frame.py --frames 4 --regime translucent --zoom 4 > gpurun_out/r4_ncu_t.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:init_rays --launch-skip 3 -c 1 -f -o gpurun_out/r4_init_rays python tools/profile_frame.py --frames 4 > gpurun_out/r4_ncu_i.log 2>&1
timeout 900 python tools/run_configs.py > gpurun_out/r4_configs.jsonl 2> gpurun_out/r4_configs.err; tail -c 400 gpurun_out/r4_configs.jsonl
ls -la gpurun_out/r4_*

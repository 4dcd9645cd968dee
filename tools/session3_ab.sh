# A/B of the set-up walk variants (session 3): direct landing (default build) vs the plain loop, walk batches of 2 / 8
set -x
( time timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_probes.py tests/test_gpu_vs_reference.py -q -x -k "traversal or first_hit or aabb_scale_4 or probe" ) > gpurun_out/s3_ab_pytest.log 2>&1; tail -3 gpurun_out/s3_ab_pytest.log
for rep in 1 2; do
for lib in libnmr.so libnmr_loop.so libnmr_wb8.so libnmr_wb2.so; do
  bash tools/ab1.sh "rep$rep $lib" nerf-glasses_b200/$lib
  bash tools/ab1.sh "rep$rep $lib zoom4" nerf-glasses_b200/$lib --zoom 4
done
done > gpurun_out/s3_ab.txt 2>&1
cat gpurun_out/s3_ab.txt

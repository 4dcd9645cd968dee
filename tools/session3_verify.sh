set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/s3_pytest.log 2>&1; tail -4 gpurun_out/s3_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke.log 2>&1; tail -2 gpurun_out/s3_smoke.log
( time timeout 500 python bench.py ) > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; tail -c 300 gpurun_out/s3_bench.json; tail -3 gpurun_out/s3_bench.err

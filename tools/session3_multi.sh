# 2-GPU evidence on the final build: the multi-GPU tests the driver's one-GPU box skips, and the N=2 bench line
set -x
nvidia-smi -L
( time timeout 300 python -m pytest tests/test_gpu_multi.py -q -x ) > gpurun_out/s3_multi_pytest.log 2>&1; tail -4 gpurun_out/s3_multi_pytest.log
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 ) > gpurun_out/s3_bench_n2.json 2> gpurun_out/s3_bench_n2.err; tail -c 400 gpurun_out/s3_bench_n2.json; tail -3 gpurun_out/s3_bench_n2.err

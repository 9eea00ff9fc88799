mkdir -p gpurun_out
N=2
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r2b_multigpu_tests_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2b_bench_cfg4_n$N.json 2> gpurun_out/r2b_bench_cfg4_n$N.err
tail -c 1200 gpurun_out/r2b_bench_cfg4_n$N.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --workload cfg5 --steps 2 --warmup 2 > gpurun_out/r2b_bench_cfg5_n$N.json 2> gpurun_out/r2b_bench_cfg5_n$N.err
tail -c 500 gpurun_out/r2b_bench_cfg5_n$N.json; echo; tail -3 gpurun_out/r2b_bench_cfg4_n$N.err

# usage: bash tools/gpu_scale.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r2z_multigpu_tests_n$N.log
if [ "$N" = "1" ]; then
  python bench.py --steps 3 --warmup 3 > gpurun_out/r2z_bench_cfg4_n1.json 2> gpurun_out/r2z_bench_cfg4_n1.err
  python bench.py --workload cfg5 --steps 2 --warmup 1 > gpurun_out/r2z_bench_cfg5_n1.json 2> gpurun_out/r2z_bench_cfg5_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2z_bench_cfg4_n$N.json 2> gpurun_out/r2z_bench_cfg4_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --workload cfg5 --steps 2 --warmup 1 > gpurun_out/r2z_bench_cfg5_n$N.json 2> gpurun_out/r2z_bench_cfg5_n$N.err
fi
tail -c 600 gpurun_out/r2z_bench_cfg4_n$N.json; echo; tail -c 400 gpurun_out/r2z_bench_cfg5_n$N.json; echo; tail -3 gpurun_out/r2z_bench_cfg4_n$N.err

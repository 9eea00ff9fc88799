mkdir -p gpurun_out
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --workload cfg5 --steps 2 --warmup 2 > gpurun_out/r2b_bench_cfg5_n$N.json 2> gpurun_out/r2b_bench_cfg5_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_cfg5_n2.json').read().strip().splitlines()[-1]); print('cfg5 N=2 value', round(d['value'],1), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],1))"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/r2b_multigpu_tests_n$N.log

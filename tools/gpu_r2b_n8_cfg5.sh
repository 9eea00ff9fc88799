# usage: bash tools/gpu_r2b_n8_cfg5.sh N   (under gpurun --gpus N): cfg 5 with the overlapped feature download
N=$1
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --workload cfg5 --steps 2 --warmup 2 > gpurun_out/r2b_bench_cfg5_n$N.json 2> gpurun_out/r2b_bench_cfg5_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_cfg5_n$N.json').read().strip().splitlines()[-1]); print('cfg5 N=$N value', round(d['value'],1), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],1))"

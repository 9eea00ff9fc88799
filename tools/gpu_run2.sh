mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "multi or sharded or two_devices or cfg12 or match_api or match_post_opt or knn_search_vs or weight_fit_matches" 2>&1 | tail -40 > gpurun_out/r2b_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/r2b_bench_n2_p2p.json 2> gpurun_out/r2b_bench_n2_p2p.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 --exchange reduce_scatter > gpurun_out/r2b_bench_n2_rs.json 2> gpurun_out/r2b_bench_n2_rs.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --workload cfg5 > gpurun_out/r2b_bench_cfg5_n2.json 2> gpurun_out/r2b_bench_cfg5_n2.err
tail -8 gpurun_out/r2b_tests.log; tail -3 gpurun_out/r2b_bench_n2_p2p.err

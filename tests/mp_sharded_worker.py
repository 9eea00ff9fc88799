"""Worker of tests/test_multi_gpu.py — run as
    python -m torch.distributed.run --nproc-per-node N tests/mp_sharded_worker.py
one rank per GPU over NCCL.  Every rank holds the WHOLE (small) pool so that it can compare the
sharded path (its shard + NCCL merge + peer-memory gather) with one single-GPU search of the
concatenated pool on the same inputs: distances, indices and matched features must be
bit-identical, ties across the shard boundary included."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from knn_svc_b200 import ops, sharded, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    checks = 0
    for case, (T, NP, D) in {"dense": (701, 20011, 1024), "randn": (300, 9000, 1024), "odd_dim": (130, 5003, 192)}.items():
        if case == "dense":
            pool = synth.ar1_frames(NP, seed=11)
            query = synth.ar1_frames(T, seed=12)
        else:
            pool = synth.randn_frames(NP, d=D, seed=13) + 0.25
            query = synth.randn_frames(T, d=D, seed=14) + 0.25
        # exact ties that straddle every shard boundary, and queries that sit on them
        for r in range(1, world):
            b = sharded.shard_bounds(NP, world, r)[0]
            pool[b] = pool[b - 1]
            pool[b + 3] = pool[b - 1]
            pool[b - 7] = pool[b - 1]
            query[r] = pool[b - 1]
            query[T - r] = pool[b - 1] * 1.5
        pool_t, query_t = torch.from_numpy(pool).to(dev), torch.from_numpy(query).to(dev)
        qp, pp = ops.prepare_rows(query_t), ops.prepare_rows(pool_t)
        lo, hi = sharded.shard_bounds(NP, world, rank)
        for exchange in ("p2p", "reduce_scatter"):
            sp = sharded.ShardedPool(pool_t[lo:hi].clone(), lo, exchange=exchange)
            assert sp.bounds[0] == 0 and sp.bounds[-1] == NP and len(sp.bounds) == world + 1
            for k in (4, 32):
                d1, i1, d1_64 = ops.knn_search(qp, pp, k, return_dist64=True)          # one GPU, whole pool
                f1 = ops.gather_mix(pp.rows, i1, None)
                m = sp.match(qp, k, gather="all")
                assert torch.equal(m.idx, i1), (case, exchange, k, "idx", int((m.idx != i1).sum()))
                assert torch.equal(m.dist, d1), (case, exchange, k, "dist")
                assert torch.equal(m.dist64, d1_64), (case, exchange, k, "dist64")
                if exchange == "p2p":
                    assert torch.equal(m.feats, f1), (case, k, "feats", float((m.feats - f1).abs().max()))
                else:
                    assert torch.allclose(m.feats, f1, rtol=0, atol=2e-6 * float(f1.abs().max())), (case, k, "feats rs")
                s = sp.match(qp, k, gather="slice")
                a, b = s.rows
                assert (a, b) == sharded.query_slice(T, world, rank)
                if exchange == "p2p":
                    assert torch.equal(s.feats, f1[a:b])
                w = torch.softmax(torch.randn((T, k), device=dev, generator=torch.Generator(device=dev).manual_seed(5)), 1)
                mw = sp.match(qp, k, gather="all", weights=w)
                fw = ops.gather_mix(pp.rows, i1, w)
                if exchange == "p2p":
                    assert torch.equal(mw.feats, fw), (case, k, "weighted feats")
                checks += 1
            sp.close()
    # the matcher API on a sharded pool (KNeighborsVC.match): every rank gets all rows
    from knn_svc_b200.ddsp_matcher import KNeighborsVC
    knn = KNeighborsVC(None, None, None, device=dev)
    pool = synth.ar1_frames(6000, seed=21)
    query = synth.ar1_frames(257, seed=22)
    pool_t, query_t = torch.from_numpy(pool).to(dev), torch.from_numpy(query).to(dev)
    lo, hi = sharded.shard_bounds(6000, world, rank)
    sp = sharded.ShardedPool(pool_t[lo:hi].clone(), lo)
    got = knn.match(query_t, sp, topk=4, without_vocode=True)
    want = knn.match(query_t, pool_t, topk=4, without_vocode=True)
    assert torch.equal(got, want)
    # post_opt on the sharded pool: greedy re-selection and weight fit read the pool through the peer row
    # table (`previous selection + 1` and idx +- 1 cross shard boundaries); bit-identical to one GPU
    for post_opt in ("post_opt_0.2", "post_opt_plain"):          # the second parses to -1: fit only, no re-selection
        got = knn.match(query_t, sp, topk=4, without_vocode=True, post_opt=post_opt)
        want = knn.match(query_t, pool_t, topk=4, without_vocode=True, post_opt=post_opt)
        assert torch.equal(got, want), post_opt
        checks += 1
    sp.close()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.write(f"rank{rank}-ok checks={checks}\n")
    sys.stdout.flush()


if __name__ == "__main__":
    main()

"""CPU-side checks: the C-ABI library loads and exports every declared symbol,
the host logic of the sharded exchange works over gloo with world_size 2, and
the product refuses to run without CUDA (no CPU fallback)."""
import os
import re
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from knn_svc_b200 import _lib
    header = (ROOT / "include" / "knnsvc_b200.h").read_text()
    declared = set(re.findall(r"\b(knnsvc_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.knnsvc_version() >= 100
    # pure host-side query, no GPU needed
    assert lib.knnsvc_weight_fit_workspace_bytes(100, 4) > 0


def test_filter_traversal_plan_invariants():
    """knnsvc_knn_plan is pure host logic: chains (query tile x pool segment) cut into L2-sized blocks.
    Every shape must give a plan that covers the problem, fits the grid and matches the workspace query."""
    import ctypes
    from knn_svc_b200 import _lib
    lib = _lib.load()
    shapes = [(1, 33, 4), (20, 257, 32), (3000, 30000, 4), (3001, 3001, 32), (100_000, 30_000, 32), (1500, 30_000, 4),
              (100_000, 1_250_000, 4), (100_000, 10_000_000, 4), (3000, 10_000_000, 32), (262_144, 262_144, 32),
              (129, 513, 8), (148 * 128, 256, 1), (148 * 128 + 1, 24_577, 4)]
    for T, NP, k in shapes:
        out = (ctypes.c_int * 8)()
        _lib.check(lib.knnsvc_knn_plan(T, NP, k, ctypes.cast(out, ctypes.c_void_p)), "knn_plan")
        ctas, n_qt, n_pt, n_seg, n_blk, units, grid, cap = list(out)
        assert ctas == 1 and n_qt == -(-T // 128) and n_pt == -(-NP // 256)
        assert 1 <= n_seg <= min(16, n_pt) and n_blk >= 1 and units == n_qt * n_seg * n_blk
        assert 1 <= grid <= 148 and grid == min(units, 148)
        assert cap == 2048
        seg_tiles = -(-n_pt // n_seg)
        assert -(-seg_tiles // n_blk) <= 96, "a block holds at most 96 pool tiles (48 MB of fp16 operand)"
        if n_qt >= 148:
            assert n_seg <= 12, "many query tiles: segments are not needed to fill the SMs"
        if n_qt * n_seg < 148:
            assert n_blk == 1 or n_qt * 16 < 148, "few chains: blocks of a chain would serialise"
        ws = lib.knnsvc_knn_workspace_bytes(T, NP, 1024, k)
        # candidate log (value, column) + refined value per slot, per-slot block offsets (<= 1025 ints), small terms
        assert ws >= T * n_seg * cap * 12
        assert ws < T * n_seg * cap * 12 + T * n_seg * (k + 3 + 1026) * 4 + (1 << 26)
    # the headline shape: one segment, 407 blocks of 96 tiles per chain
    out = (ctypes.c_int * 8)()
    lib.knnsvc_knn_plan(100_000, 10_000_000, 4, ctypes.cast(out, ctypes.c_void_p))
    assert list(out)[3:6] == [1, 407, 782 * 407]
    with pytest.raises(ValueError):
        _lib.check(lib.knnsvc_knn_plan(0, 10, 4, ctypes.cast(out, ctypes.c_void_p)), "knn_plan")


def test_ops_refuse_cpu_tensors():
    from knn_svc_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.prepare_rows(torch.randn(4, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.cosine_dist(torch.randn(4, 64), torch.randn(5, 64))


def test_product_does_not_import_oracle():
    for f in (ROOT / "knn_svc_b200").glob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f


def test_shard_bounds_cover_pool():
    from knn_svc_b200.sharded import shard_bounds
    for n, w in ((10_000_000, 8), (1001, 3), (7, 8)):
        edges = [shard_bounds(n, w, r) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))


def test_merge_rule_host_is_shard_invariant():
    from knn_svc_b200.sharded import merge_topk_host, shard_bounds
    rs = np.random.RandomState(0)
    d = rs.rand(17, 400).astype(np.float32)
    d[:, 100:120] = d[:, 50:51]            # exact ties across what will be different shards
    full = torch.from_numpy(d)
    k = 8
    order = torch.argsort(full, dim=1, stable=True)[:, :k]
    for world in (1, 2, 3, 8):
        gd, gi = [], []
        for r in range(world):
            lo, hi = shard_bounds(400, world, r)
            o = torch.argsort(full[:, lo:hi], dim=1, stable=True)[:, :k]
            gd.append(torch.gather(full[:, lo:hi], 1, o)); gi.append(o + lo)
        md, mi = merge_topk_host(torch.stack(gd), torch.stack(gi))
        assert torch.equal(mi, order)
        assert torch.equal(md, torch.gather(full, 1, order))


def test_query_slices_tile_the_batch():
    """the rows each rank produces matched features for: contiguous, equal-sized chunks that tile [0, T)"""
    from knn_svc_b200.sharded import query_slice
    for T in (0, 1, 7, 100, 100_000, 100_003):
        for world in (1, 2, 3, 8):
            edges = [query_slice(T, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == T
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            chunk = (T + world - 1) // world
            assert all(b - a <= chunk for a, b in edges)


WORKER = r"""
import os, sys, torch, numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
from knn_svc_b200.sharded import all_gather_topk, merge_topk_host, shard_bounds
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rs = np.random.RandomState(0)
full = torch.from_numpy(rs.rand(23, 501).astype(np.float32))
k = 4
lo, hi = shard_bounds(501, world, rank)
o = torch.argsort(full[:, lo:hi], dim=1, stable=True)[:, :k]
ld, li = torch.gather(full[:, lo:hi], 1, o), o + lo
gd, gi = all_gather_topk(ld, li)
md, mi = merge_topk_host(gd, gi)
want = torch.argsort(full, dim=1, stable=True)[:, :k]
assert torch.equal(mi, want), (rank, mi[0], want[0])
assert torch.equal(md, torch.gather(full, 1, want))
# the sharded search exchanges fp64 distances: same payload packing, dtype preserved bit for bit
gd64, gi64 = all_gather_topk(ld.double() * (1 + 2.0 ** -40), li)
assert gd64.dtype == torch.float64 and torch.equal(gi64, gi)
assert torch.equal(gd64[rank], ld.double() * (1 + 2.0 ** -40))
md64, mi64 = merge_topk_host(gd64, gi64)
assert torch.equal(mi64, want)
dist.barrier()
dist.destroy_process_group()
sys.stdout.write(f"rank{rank}-ok\n"); sys.stdout.flush()
"""


def test_sharded_exchange_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, REPO_ROOT=str(ROOT), CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("-ok") == 2, r.stdout


def test_bench_reference_arm_runs_on_host_cores():
    """`bench.py --impl reference` times the reference's CPU matcher (the oracle port) on the host cores and
    prints one JSON line with the contract's keys; it never touches the GPU library."""
    import json
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-queries", "8", "--cpu-pool", "3000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "query frames/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["config"]["workload"].startswith("cfg4")


def test_cfg5_pairs_are_dealt_completely_and_evenly():
    """bench.py --workload cfg5: every (utterance, target) pair of the split's shape goes to exactly one rank,
    ranks get equal frame counts (within one utterance) and at most a few target pools each"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for world in (1, 2, 4, 8):
        seen, frames = set(), []
        for rank in range(world):
            mine, total_pairs, total_frames = bench.cfg5_jobs(world, rank)
            assert total_pairs == 7038
            for t, jobs in mine.items():
                for su, n in jobs:
                    assert (t, su) not in seen
                    seen.add((t, su))
            frames.append(sum(n for jobs in mine.values() for _, n in jobs))
            assert len(mine) <= 3 or world == 1
        assert len(seen) == 7038 and sum(frames) == total_frames
        assert max(frames) - min(frames) <= 2 * 1500


def test_options_from_the_environment():
    """KNNSVC_OPTIONS="name=value,..." is applied when the library is loaded (no GPU needed: the switches are
    host-side atomics); an unknown name fails loudly; every switch the header documents is accepted"""
    import subprocess
    import sys
    root = Path(__file__).resolve().parent.parent
    code = "from knn_svc_b200 import _lib; _lib.load(); print('loaded')"
    ok = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True,
                        env=dict(os.environ, KNNSVC_OPTIONS="concat_cluster=0, concat_f0_table=0,refine_min_candidates=800"))
    assert ok.returncode == 0 and "loaded" in ok.stdout, ok.stderr[-2000:]
    bad = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True,
                         env=dict(os.environ, KNNSVC_OPTIONS="no_such_switch=1"))
    assert bad.returncode != 0 and "unknown option" in bad.stderr
    from knn_svc_b200 import _lib
    lib = _lib.load()
    header = (root / "include" / "knnsvc_b200.h").read_text()
    for name, value in (("concat_staged", 1), ("concat_cluster", 1), ("concat_f0_table", 1), ("weight_fit_cluster", 1),
                        ("refine_min_candidates", 0), ("log_cap", 0), ("block_tiles", 0), ("query_group", 0),
                        ("filter_flags", 1), ("spin_sleep_ns", 40), ("cta_group", 1)):
        assert f'"{name}"' in header, f"{name} is not documented in the header"
        assert lib.knnsvc_set_option(name.encode(), value) == 0, name
    assert lib.knnsvc_set_option(b"cta_group", 2) != 0       # the removed variant is refused, not ignored

"""Tiny multi-block search for compute-sanitizer runs of the filter's unit queue and state hand-over:
compute-sanitizer --tool racecheck python tools/racecheck_filter.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knn_svc_b200 import _lib, ops
lib = _lib.load()
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(0)
q = torch.randn((300, 1024), device=dev, generator=g) + 1.0
p = torch.randn((6000, 1024), device=dev, generator=g) + 1.0
qp, pp = ops.prepare_rows(q), ops.prepare_rows(p)
ref = ops.knn_exact(qp, pp, 4)
for blk, flags in ((1, 1), (1, 5), (0, 1)):
    lib.knnsvc_set_option(b"block_tiles", blk); lib.knnsvc_set_option(b"filter_flags", flags)
    d, i, st = ops.knn_search(qp, pp, 4, return_stats=True)
    torch.cuda.synchronize()
    print("block_tiles", blk, "flags", flags, "units", int(st[4]), "n_seg", int(st[3]), "same as exact:", bool((i == ref[1]).all()), flush=True)

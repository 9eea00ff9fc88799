"""A/B of the filter's unit order: chains per group x pool tiles per block -> filter time (CUDA events,
plain run) and, when run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`, the DRAM
bytes of each filter launch (launches appear in the order printed here).
    python tools/traffic_ab.py [T NP]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knn_svc_b200 import _lib, ops

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 2500000
reps = int(os.environ.get("REPS", 2))
dev = "cuda:0"
lib = _lib.load()
g = torch.Generator(device=dev); g.manual_seed(0)
q = torch.randn((T, 1024), device=dev, generator=g)
p = torch.empty((NP, 1024), device=dev)
for a in range(0, NP, 1 << 20):
    b = min(NP, a + (1 << 20)); p[a:b] = torch.randn((b - a, 1024), device=dev, generator=g)
qp, pp = ops.prepare_rows(q, check=False), ops.prepare_rows(p, check=False)
ref = None
CONFIGS = ((1 << 20, 96, 1), (148, 96, 1), (296, 96, 1), (296, 48, 1), (444, 48, 1), (592, 32, 1))
if os.environ.get("CONFIGS") == "default":
    CONFIGS = ((0, 0, 1),)
for grp, blk, cta in CONFIGS:
    lib.knnsvc_set_option(b"query_group", grp); lib.knnsvc_set_option(b"block_tiles", blk)
    d, i = ops.knn_search(qp, pp, 4); torch.cuda.synchronize()
    lib.knnsvc_filter_timing(1)
    for _ in range(reps):
        d, i = ops.knn_search(qp, pp, 4)
    torch.cuda.synchronize()
    buf = (ctypes.c_float * 256)()
    n = lib.knnsvc_filter_timing_collect(ctypes.cast(buf, ctypes.c_void_p), 256)
    lib.knnsvc_filter_timing(0)
    ms = sum(buf[j] for j in range(n)) / max(n, 1)
    same = True if ref is None else bool((i == ref[1]).all() and (d == ref[0]).all())
    if ref is None:
        ref = (d.clone(), i.clone())
    print(json.dumps({"T": T, "NP": NP, "query_group": grp, "block_tiles": blk, "filter_ms": round(ms, 3),
                      "TFLOPs": round(2.0 * T * NP * 1024 / ms / 1e9, 1), "launches": reps + 1, "identical_results": same}), flush=True)
lib.knnsvc_set_option(b"query_group", 0); lib.knnsvc_set_option(b"block_tiles", 0)

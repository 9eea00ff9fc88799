"""Sweep of the filter's L2 block size (pool tiles per block) under sustained load, alternating in one process.
BLOCKS=16,32,48,64,96,1000000 (a huge value = one block per chain, the pre-blocking traversal)."""
import ctypes, os, statistics, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knn_svc_b200 import _lib, ops

lib = _lib.load()
dev = "cuda:0"
T, NP, K = int(os.environ.get("T", 100000)), int(os.environ.get("NP", 10000000)), int(os.environ.get("K", 4))
g = torch.Generator(device=dev); g.manual_seed(0)
q = torch.randn((T, 1024), device=dev, generator=g)
p = torch.empty((NP, 1024), device=dev)
for a in range(0, NP, 1 << 20):
    b = min(NP, a + (1 << 20)); p[a:b] = torch.randn((b - a, 1024), device=dev, generator=g)
qp, pp = ops.prepare_rows(q, check=False), ops.prepare_rows(p, check=False)


def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                          capture_output=True, text=True).stdout.strip()


ref = None
# CONFIGS = block_tiles:filter_flags pairs (flags bit0: query-tile prefetch, bit1: pool-block prefetch, bit2: static schedule)
configs = [tuple(int(v) for v in c.split(":")) for c in
           os.environ.get("CONFIGS", "1000000:4,1000000:0,48:4,48:0,48:1,48:3,32:3,64:3,96:3").split(",")]
for rep in range(int(os.environ.get("REPS", 2))):
    for blk, flags in configs:
        lib.knnsvc_set_option(b"block_tiles", blk); lib.knnsvc_set_option(b"filter_flags", flags)
        d, i, st = ops.knn_search(qp, pp, K, return_stats=True); torch.cuda.synchronize()
        if ref is None:
            ref = (d.clone(), i.clone())
        same = bool((i == ref[1]).all()) and bool((d == ref[0]).all())
        lib.knnsvc_filter_timing(1)
        samples, stop = [], [False]
        def samp():
            while not stop[0]:
                samples.append(clocks()); time.sleep(0.25)
        th = threading.Thread(target=samp); th.start()
        steps = int(os.environ.get("STEPS", 3))
        for _ in range(steps):
            d, i, st = ops.knn_search(qp, pp, K, return_stats=True)
        torch.cuda.synchronize()
        stop[0] = True; th.join()
        buf = (ctypes.c_float * 256)()
        n = lib.knnsvc_filter_timing_collect(ctypes.cast(buf, ctypes.c_void_p), 256)
        lib.knnsvc_filter_timing(0)
        ms = sum(buf[j] for j in range(n)) / n
        mhz = [float(s.split(",")[0]) for s in samples if s] or [0]
        print(f"rep{rep} block_tiles {blk:8d} flags {flags}: filter {ms:8.2f} ms {2.0*T*NP*1024/ms/1e9:7.1f} TF  clk {statistics.median(mhz):.0f}  "
              f"n_seg {int(st[3])} units {int(st[4])} logged/row {int(st[1])/T:.1f} surv/row {int(st[2])/T:.1f} flagged {int(st[0])} same-as-first {same}",
              flush=True)
lib.knnsvc_set_option(b"block_tiles", 0); lib.knnsvc_set_option(b"filter_flags", 3)

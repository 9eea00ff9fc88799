"""A/B of the filter's operand formats under sustained, power-capped load: fp16 vs bf16.  Reports filter TFLOP/s, clocks, logged candidates and
survivors per row — on randn rows (BASELINE cfg 4 shape, smaller pool) and on WavLM-like AR(1) rows."""
import ctypes, os, statistics, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knn_svc_b200 import _lib, ops, synth

lib = _lib.load()
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(0)


def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                          capture_output=True, text=True).stdout.strip()


def prep(x, bf):
    lib.knnsvc_set_option(b"bf16_operands", bf)
    return ops.prepare_rows(x, check=False)


def run(tag, q, p, k, fq, fp, steps):
    qp, pp = prep(q, fq), prep(p, fp)
    assert fq == fp   # kind::f16 with A fp16 and B bf16 is an illegal instruction on sm_100a (tried)
    lib.knnsvc_set_option(b"bf16_operands", fq)
    ops.knn_search(qp, pp, k); torch.cuda.synchronize()
    lib.knnsvc_filter_timing(1)
    samples, stop = [], [False]
    def samp():
        while not stop[0]:
            samples.append(clocks()); time.sleep(0.2)
    th = threading.Thread(target=samp); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        d, i, st = ops.knn_search(qp, pp, k, return_stats=True)
    e1.record(); torch.cuda.synchronize()
    stop[0] = True; th.join()
    buf = (ctypes.c_float * 256)()
    n = lib.knnsvc_filter_timing_collect(ctypes.cast(buf, ctypes.c_void_p), 256)
    lib.knnsvc_filter_timing(0)
    lib.knnsvc_set_option(b"bf16_operands", 0)
    ms = sum(buf[j] for j in range(n)) / n
    tot = e0.elapsed_time(e1) / steps
    mhz = [float(s.split(",")[0]) for s in samples if s] or [0]
    T, NP = qp.n, pp.n
    print(f"{tag} k={k} q={'bf16' if fq else 'fp16'} p={'bf16' if fp else 'fp16'}: filter {ms:8.2f} ms {2.0*T*NP*1024/ms/1e9:7.1f} TF  "
          f"search {tot:8.2f} ms  clk {statistics.median(mhz):.0f}  err q {float(qp.err):.2e} p {float(pp.err):.2e}  "
          f"flagged {int(st[0])} logged/row {int(st[1])/T:.1f} surv/row {int(st[2])/T:.1f}", flush=True)
    return d, i


def ar1(n, seed):
    base = torch.from_numpy(synth.ar1_frames(min(n, 6000), seed=seed)).to(dev)
    x = base.repeat((n + len(base) - 1) // len(base), 1)[:n].contiguous()
    return x + 0.05 * torch.randn(x.shape, device=dev, generator=g)


FMTS = ((0, 0), (1, 1))   # mixed fp16 x bf16 is an illegal instruction for kind::f16 (tried)
T, NP = int(os.environ.get("T", 100000)), int(os.environ.get("NP", 4000000))
q = torch.randn((T, 1024), device=dev, generator=g)
p = torch.empty((NP, 1024), device=dev)
for a in range(0, NP, 1 << 20):
    b = min(NP, a + (1 << 20)); p[a:b] = torch.randn((b - a, 1024), device=dev, generator=g)
ref = None
for rep in range(int(os.environ.get("REPS", 2))):
    for fq, fp in FMTS:
        d, i = run(f"randn {T}x{NP} rep{rep}", q, p, 4, fq, fp, 3)
        if ref is None:
            ref = (d.clone(), i.clone())
        else:
            print("   identical to fp16/fp16:", bool((i == ref[1]).all()), bool((d == ref[0]).all()), flush=True)
d, i = run(f"randn {T}x{NP}", q, p, 32, 0, 0, 1); ref = (d.clone(), i.clone())
for fq, fp in FMTS[1:]:
    d, i = run(f"randn {T}x{NP}", q, p, 32, fq, fp, 1)
    print("   identical to fp16/fp16:", bool((i == ref[1]).all()), bool((d == ref[0]).all()), flush=True)
del p, q
for (T2, NP2) in ((20000, 1000000), (100000, 30000)):
    q, p = ar1(T2, 1), ar1(NP2, 2)
    for k in (4, 32):
        ref = None
        for fq, fp in FMTS:
            d, i = run(f"ar1 {T2}x{NP2}", q, p, k, fq, fp, 2)
            if ref is None:
                ref = (d.clone(), i.clone())
            else:
                print("   identical to fp16/fp16:", bool((i == ref[1]).all()), bool((d == ref[0]).all()), flush=True)

"""A/B of filter-kernel variants under sustained (power-capped) load, alternating in one process."""
import sys, os, time, ctypes, subprocess, threading, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knn_svc_b200 import ops, _lib

lib = _lib.load()
dev = "cuda:0"
T, NP = int(os.environ.get("T", 100000)), int(os.environ.get("NP", 4000000))
g = torch.Generator(device=dev); g.manual_seed(0)
q = torch.randn((T, 1024), device=dev, generator=g)
p = torch.empty((NP, 1024), device=dev)
for a in range(0, NP, 1 << 20):
    b = min(NP, a + (1 << 20)); p[a:b] = torch.randn((b - a, 1024), device=dev, generator=g)

def clocks():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True).stdout.strip()
    return out

K = int(os.environ.get("K", 4))
configs = [("cta1 fp16", 1, 0), ("cta1 bf16", 1, 1)]      # (the cta_group::2 variant was removed in round 2)
if os.environ.get("EPI"):      # sweep the epilogue poll back-off instead: EPI=0,100,400
    configs = [(f"cta1 fp16 epi{e}", 1, 0, int(e)) for e in os.environ["EPI"].split(",")]
if os.environ.get("CONFIGS"):
    configs = [c for c in configs if c[0] in os.environ["CONFIGS"].split(",")]
reps = int(os.environ.get("REPS", 3))
steps = int(os.environ.get("STEPS", 4))
for rep in range(reps):
    for cfg in configs:
        name, cg, bf = cfg[:3]
        lib.knnsvc_set_option(b"bf16_operands", bf)
        lib.knnsvc_set_option(b"epi_sleep_ns", cfg[3] if len(cfg) > 3 else 0)
        lib.knnsvc_set_option(b"spin_sleep_ns", int(os.environ.get("SPIN", 0)))
        qp, pp = ops.prepare_rows(q, check=False), ops.prepare_rows(p, check=False)
        ops.knn_search(qp, pp, K); torch.cuda.synchronize()
        lib.knnsvc_filter_timing(1)
        samples = []
        stop = False
        def samp():
            while not stop:
                samples.append(clocks()); time.sleep(0.25)
        th = threading.Thread(target=samp); th.start()
        t0 = time.time()
        for _ in range(steps):
            d, i, st = ops.knn_search(qp, pp, K, return_stats=True)
        torch.cuda.synchronize()
        wall = (time.time() - t0) / steps * 1e3
        stop = True; th.join()
        buf = (ctypes.c_float * 256)()
        n = lib.knnsvc_filter_timing_collect(ctypes.cast(buf, ctypes.c_void_p), 256)
        lib.knnsvc_filter_timing(0)
        ms = sum(buf[j] for j in range(n)) / n
        mhz = [float(s.split(",")[0]) for s in samples if s]
        pw = [float(s.split(",")[1]) for s in samples if s]
        print(f"rep{rep} {name}: filter {ms:8.2f} ms  {2.0*T*NP*1024/ms/1e9:7.1f} TFLOP/s  clk med {statistics.median(mhz):.0f} MHz  pw med {statistics.median(pw):.0f} W"
              f"  temp {samples[-1].split(',')[2] if samples else '?'}  flagged {int(st[0])} logged {int(st[1])} survivors {int(st[2])} search wall {wall:.1f} ms", flush=True)

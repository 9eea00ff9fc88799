"""Does the download of a batch's matched features overlap the next batch's matching?  (cfg-5 e2e leg)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from knn_svc_b200 import synth
from knn_svc_b200 import ddsp_prematch_dataset as pm
import bench
dev = torch.device("cuda:0")


class Env:
    pass


env = Env(); env.torch = torch; env.dev = dev
pool = bench._pipeline_pool(env, 30000, 900)
rs = np.random.RandomState(0)
lens = rs.randint(150, 1501, size=384)
utts = [(synth.ar1_frames_device(int(n), 1024, seed=100 + i, device=dev, seg_len=200), torch.from_numpy(synth.f0_track(int(n), seed=500 + i)))
        for i, n in enumerate(lens)]
total = int(lens.sum())
host_small = torch.empty((1500, 1024)).pin_memory()
host_big = torch.empty((total, 1024)).pin_memory()
copy_stream = torch.cuda.Stream(device=dev)
NB = 4


def run(mode):
    main = torch.cuda.current_stream(dev)
    for b in range(NB):
        res = pm.match_utterances([u[0] for u in utts], [u[1] for u in utts], pool, post_opt="post_opt_0.2", ckpt_type="mix",
                                  prioritize_f0=True)
        if mode == "per_utt":
            copy_stream.wait_stream(main)
            with torch.cuda.stream(copy_stream):
                for r in res:
                    host_small[:r["out_feats"].shape[0]].copy_(r["out_feats"], non_blocking=True)
                    r["out_feats"].record_stream(copy_stream)
        elif mode == "per_utt_main":
            for r in res:
                host_small[:r["out_feats"].shape[0]].copy_(r["out_feats"], non_blocking=True)
        elif mode == "one_copy":
            base = res[0]["out_feats"]._base
            copy_stream.wait_stream(main)
            with torch.cuda.stream(copy_stream):
                host_big[:base.shape[0]].copy_(base, non_blocking=True)
                base.record_stream(copy_stream)
    if mode in ("per_utt", "one_copy"):
        main.wait_stream(copy_stream)


for mode in ("none", "per_utt_main", "per_utt", "one_copy", "none"):
    run(mode); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record(); run(mode); e1.record(); torch.cuda.synchronize()
    print(f"{mode:14s} {e0.elapsed_time(e1):8.1f} ms device, {1e3 * (time.time() - t0):8.1f} ms wall; {NB} batches of {len(utts)} utterances, "
          f"{total} frames = {total * 4096 / 1e9:.2f} GB of features each", flush=True)
x = torch.empty((total, 1024), device=dev)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); host_big.copy_(x, non_blocking=True); e1.record(); torch.cuda.synchronize()
print(f"D2H alone: {e0.elapsed_time(e1):.1f} ms for {total * 4096 / 1e9:.2f} GB = {total * 4096 / 1e6 / e0.elapsed_time(e1):.1f} GB/s")

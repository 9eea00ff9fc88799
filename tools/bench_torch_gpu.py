"""Context number for DESIGN.md: the reference's matcher arithmetic executed by stock PyTorch ON THE
SAME B200 (the reference ships no CUDA of its own) next to this library, at BASELINE cfg 3
(3000 x 30000, k=32 as the live path takes, ddsp_prematch_dataset.py:1196-1206) and a slice of cfg 4.
 (a) the reference's own loop shape: 20-row chunks, torch.cdist + norms + topk per chunk;
 (b) the best one can do with library calls: one fp32 (TF32 off) matmul per 4096-row block + topk.
Not a product path and not the bench's baseline (that is the CPU arm)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knn_svc_b200 import ops

dev = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device=dev); g.manual_seed(0)


def ref_loop(q, p, k=32, inc=20):
    pn = torch.norm(p, p=2, dim=-1)
    out = []
    for a in range(0, len(q), inc):
        s = q[a:a + inc]
        sn = torch.norm(s, p=2, dim=-1)
        d = torch.cdist(s[None], p[None], p=2)[0]
        dot = (-(d ** 2) + sn[:, None] ** 2 + pn[None] ** 2) / 2
        out.append((1 - dot / (sn[:, None] * pn[None])).topk(k=k, dim=-1, largest=False).indices)
    return torch.cat(out)


def blocked(q, p, k=32, blk=4096):
    pu = p / torch.norm(p, dim=-1, keepdim=True)
    out = []
    for a in range(0, len(q), blk):
        s = q[a:a + blk]
        su = s / torch.norm(s, dim=-1, keepdim=True)
        out.append((1 - su @ pu.T).topk(k=k, dim=-1, largest=False).indices)
    return torch.cat(out)


def ours(q, p, k=32):
    return ops.knn_search(ops.prepare_rows(q, check=False), ops.prepare_rows(p, check=False), k)[1]


def timed(fn, *a, reps=3):
    fn(*a); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn(*a)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


for (T, NP, k) in ((3000, 30_000, 32), (3000, 30_000, 4), (20_000, 2_000_000, 4)):
    q = torch.randn((T, 1024), device=dev, generator=g); p = torch.randn((NP, 1024), device=dev, generator=g)
    row = {"T": T, "Np": NP, "k": k}
    ms_o, i_o = timed(ours, q, p, k)
    row["this_library_ms"] = round(ms_o, 3)
    if T * NP <= 1e9:
        ms_r, i_r = timed(ref_loop, q, p, k)
        row["torch_reference_loop_ms"] = round(ms_r, 3)
        row["agree_with_reference_loop"] = round((i_r == i_o).float().mean().item(), 5)
    ms_b, i_b = timed(blocked, q, p, k, 4096 if NP <= 100_000 else 512)
    row["torch_blocked_fp32_matmul_ms"] = round(ms_b, 3)
    row["agree_with_blocked"] = round((i_b == i_o).float().mean().item(), 5)
    row["query_frames_per_s"] = {"this_library": round(T / ms_o * 1e3), "torch_blocked": round(T / ms_b * 1e3)}
    print(json.dumps(row), flush=True)

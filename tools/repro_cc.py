import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knn_svc_b200 import ops, synth
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(0)
n_utt, t_utt, n_pool = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
Tq = n_utt * t_utt
q = torch.from_numpy(synth.ar1_frames(min(Tq, 4000), seed=1)).to(dev).repeat((Tq + 3999) // 4000, 1)[:Tq].contiguous()
p = torch.from_numpy(synth.ar1_frames(min(n_pool, 4000), seed=2)).to(dev).repeat((n_pool + 3999) // 4000, 1)[:n_pool].contiguous()
p = p + 0.01 * torch.randn(p.shape, device=dev, generator=g)
d, nb, st = ops.knn_search(ops.prepare_rows(q), ops.prepare_rows(p), 4, return_stats=True)
torch.cuda.synchronize(); print("knn ok", st.tolist(), nb.min().item(), nb.max().item(), flush=True)
offs = [i * t_utt for i in range(n_utt + 1)]
out = ops.concat_cost_reselect(nb, q, p, utt_offsets=offs)
torch.cuda.synchronize(); print("cc ok", out.min().item(), out.max().item(), flush=True)

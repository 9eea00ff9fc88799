"""Multi-GPU parity on hardware (needs >= 2 CUDA devices; skipped otherwise).

  * the sharded path under torch.distributed.run / NCCL returns, on every rank, bit for bit what one
    single-GPU search of the concatenated pool returns — distances, indices, matched features —
    including exact ties across the shard boundary (tests/mp_sharded_worker.py);
  * one process can use two devices in turn (per-device function attributes).

Run on a multi-GPU box:  gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu -q
(the log of that run is committed under profiles/)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _n_gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_sharded_path_equals_single_gpu_search(nproc):
    if _n_gpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        str(ROOT / "tests" / "mp_sharded_worker.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ))
    assert r.returncode == 0, (r.stdout[-3000:] + r.stderr[-6000:])
    assert r.stdout.count("-ok") == nproc, r.stdout[-3000:]


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_one_process_two_devices():
    """every kernel family with an opt-in shared-memory size runs on cuda:0 and then on cuda:1 of the
    same process and returns the same bits"""
    from knn_svc_b200 import ops, synth
    from knn_svc_b200 import ddsp_prematch_dataset as pm
    q, p = synth.ar1_frames(300, seed=1), synth.ar1_frames(3000, seed=2)
    f0q, f0p = synth.f0_track(300, seed=3), synth.f0_track(3000, seed=4)
    res = []
    for dev in ("cuda:0", "cuda:1"):
        qp, pp = ops.prepare_rows(torch.from_numpy(q).to(dev)), ops.prepare_rows(torch.from_numpy(p).to(dev))
        d, i = ops.knn_search(qp, pp, 32)                                       # tcgen05 filter (101 KB+ smem)
        de, ie = ops.knn_exact(qp, pp, 4)                                       # exact kernel (200 KB smem)
        top4 = i[:, :4].contiguous()
        sel = pm.knn_with_concat_cost(top4, qp.rows, pp.rows, torch.from_numpy(f0q).to(dev),
                                      torch.from_numpy(f0p).to(dev), concat_weight=0.2)      # staged K5 (156 KB)
        w = pm.compute_wavlm_weight(sel, pp.rows)                              # K6 (cluster or one CTA)
        res.append([t.cpu() for t in (d, i, de, ie, sel, w)])
    for a, b in zip(*res):
        assert torch.equal(a, b)

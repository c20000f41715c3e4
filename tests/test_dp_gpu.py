"""Data parallelism on hardware (needs 2 GPUs; skipped on a single-GPU box): N-GPU gradients equal 1-GPU gradients on
the same global batch (SURVEY.md section 8e "parity scope").

Every rank holds a replica, takes rows [r*B/n, (r+1)*B/n) of the global batch (the reference's get_slice,
spnet/multi_gpu.py:49-54) and the gradients are averaged by the bucketed NCCL all-reduce of multi_gpu.GradAllReduce.
Train-mode BatchNorm couples the samples of a batch (per-replica statistics are the reference's tower semantics), so
the equality is checked with every BatchNorm normalising with its moving statistics inside the training step
(engine.bn_use_moving = "all"): then the loss is a plain mean over samples and the averaged shard gradients must equal
the full-batch gradient - to fp32 rounding with fp32 communication, to bf16 rounding with bf16 buckets."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from spnet_b200 import multi_gpu
from spnet_b200.engine import XceptionSPNetEngine
from spnet_b200.selfcheck import make_case
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = "cuda:" + os.environ["LOCAL_RANK"]
torch.cuda.set_device(dev)
dist.init_process_group("nccl")
H, W, Bg = 96, 128, 8
w, x, yt = make_case(H, W, Bg, seed=61)
lo, hi = multi_gpu.batch_slice(Bg, rank, world)
report = {}
for dtype, comm, tol in (("fp32", "fp32", 2e-4), ("bf16", "fp32", 3e-2), ("bf16", "bf16", 3e-2)):
    eng = XceptionSPNetEngine(H, W, hi - lo, dtype=dtype, weights=w, dropout_rate=0.0, deterministic=True, device=dev)
    eng.bn_use_moving = "all"
    hook = multi_gpu.attach_data_parallel(eng, comm_dtype=comm)
    eng.load_batch(x[lo:hi], yt[lo:hi])
    p_before = eng.params.clone()
    eng.train_step(lr=1e-4)
    torch.cuda.synchronize()
    assert not torch.equal(eng.params, p_before)          # the per-bucket Adam ran
    g_dp = (hook.glp.float() if comm == "bf16" else eng.grads) / world
    # every rank must end the step with identical weights
    chk = eng.params.double().sum().reshape(1).clone()
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    assert all(torch.equal(gathered[0], t) for t in gathered), gathered
    if rank == 0:
        one = XceptionSPNetEngine(H, W, Bg, dtype=dtype, weights=w, dropout_rate=0.0, deterministic=True, device=dev)
        one.bn_use_moving = "all"
        one.load_batch(x, yt)
        one.grad_hook = lambda e: None
        one.train_step(lr=1e-4)
        torch.cuda.synchronize()
        worst = 0.0
        gn = torch.stack([one.g[k].double().norm() / one.g[k].numel() ** 0.5 for k in one.g])
        floor = 1e-3 * float(gn.median())
        for k, (o, n, _) in one.offsets.items():
            a, b = g_dp[o:o + n].double(), one.grads[o:o + n].double()
            e = float((a - b).norm() / max(float(b.norm()), floor * n ** 0.5))
            worst = max(worst, e)
            assert e < tol, (dtype, comm, k, e)
        report[dtype + "/" + comm] = worst
        del one
    del eng, hook
    dist.barrier()
if rank == 0:
    print("DP_GRADS_OK", report)
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_gradients_equal_one_gpu_gradients(tmp_path):
    script = tmp_path / "dp_grads.py"
    script.write_text(SCRIPT % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29581", str(script)], capture_output=True, text=True, timeout=900)
    assert "DP_GRADS_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    os.makedirs(os.path.join(ROOT, "gpurun_out", "parity"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity", "dp_2gpu_gradients.txt"), "w") as f:
        f.write(out.stdout[out.stdout.index("DP_GRADS_OK"):])

"""Sharded contrastive loss with the in-kernel NVLink gather, world size 2+ (needs >= 2 GPUs on the box;
skipped on the single-GPU round-end box).  The host-side logic of the same path is covered on CPU with
gloo in tests/test_comm_cpu.py."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("b_local", [96, 300])
def test_fused_p2p_loss_equals_single_process_and_nccl(b_local):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "mgpu_worker.py"), str(b_local), "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("MGPU_RESULT ")][-1]
    out = json.loads(line[len("MGPU_RESULT "):])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"parity_mgpu_b{b_local}.json"), "w") as f:
        json.dump(out, f, indent=1)
    ref = out["single_process"]
    for it in range(3):
        # identical bf16 embeddings and identical tile arithmetic: only the reduction tree over ranks differs
        assert abs(out[f"fused_p2p_{it}"] - ref) <= 2e-6 * abs(ref), out
    assert out["fused_p2p_0"] == out["fused_p2p_1"] == out["fused_p2p_2"] == out["fused_p2p_micro"]
    assert out["grad_image_rel"] <= 1e-5 and out["grad_text_rel"] <= 1e-5, out   # sharded backward = single-process backward
    # sharded training step: all-reduced local gradients = single-process gradients of the global-batch loss.  The towers see
    # different batch shapes (b_local vs G rows), so GEMM tile / split boundaries and with them fp32 summation orders differ.
    assert abs(out["train_loss_p2p"] - out["train_loss_single"]) <= 2e-6 * abs(out["train_loss_single"]), out
    assert out["train_grad_aggregate_rel"] <= 2e-3 and out["train_grad_worst"][1] <= 2e-2, out
    assert abs(out["nccl_gather_fp64"] - ref) <= 1e-3 * abs(ref), out      # fp32 vs bf16-rounded embeddings
    assert abs(out["forward_logits_loss"] - ref) <= 1e-3 * abs(ref), out

"""Worker for tests/test_multigpu.py (launched by torch.distributed.run, one rank per GPU).

Checks the sharded contrastive step: every rank encodes its own shard, the fused loss kernel reads the
peers' embeddings over NVLink (no collective on the data path), and the result must equal
  (a) the single-process loss over the whole global batch (one handle, world = 1), and
  (b) the reference-shaped comparator: NCCL all-gather (lib/utils/comm.py:140-154) + logits + symmetric CE.
"""
import json
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from msclip_b200 import synth                      # noqa: E402
from msclip_b200.comm import gather_tensors       # noqa: E402
from msclip_b200.config import MSCLIPConfig       # noqa: E402
from msclip_b200.model import CLIP                # noqa: E402
from oracle import msclip_oracle as O             # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    b_local = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = MSCLIPConfig(layers=layers, gather_tensors=True)
    sd_np = synth.synth_state_dict(cfg, seed=5, logit_scale=math.log(1 / 0.07))
    sd = {k: torch.as_tensor(v) for k, v in sd_np.items()}
    model = CLIP(cfg)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    model.setup_data_parallel(b_local)
    G = world * b_local
    img = torch.from_numpy(synth.synth_images(b_local, 21, offset=rank * b_local)).to(dev)
    tok = torch.from_numpy(synth.synth_tokens(b_local, 21, offset=rank * b_local, ragged=True)).to(dev)
    out = {}
    for it in range(3):                               # several epochs: exercises the double-buffered exchange
        loss = float(model.contrastive_loss(img, tok))
        out[f"fused_p2p_{it}"] = loss
    # backward of the last step: this rank's gradient rows (second peer read inside the kernels)
    gi, gt = model.contrastive_loss_backward()
    # micro-batched shard (msclip_encode_pairs x n + one loss): bit-identical to the one-shot call
    out["fused_p2p_micro"] = float(model.contrastive_loss(img, tok, micro_batch=max(8, b_local // 3)))
    # (b) NCCL comparator with the same per-rank features
    fi, ft = model.encode_image(img), model.encode_text(tok)
    fi_all, ft_all = gather_tensors(fi), gather_tensors(ft)
    logits = math.exp(float(model.logit_scale)) * fi_all.double() @ ft_all.double().t()
    out["nccl_gather_fp64"] = float(O.contrastive_loss(logits))
    logits_api = model(img, tok)                      # CLIP.forward with GATHER_TENSORS -> [G, G] on every rank
    assert logits_api.shape == (G, G)
    out["forward_logits_loss"] = float(O.contrastive_loss(logits_api.double()))
    # (a) single-process truth on rank 0 over the whole global batch
    if rank == 0:
        solo = CLIP(MSCLIPConfig(layers=layers))
        solo.load_state_dict(sd)
        solo = solo.to(dev).eval()
        img_all = torch.from_numpy(synth.synth_images(G, 21)).to(dev)
        tok_all = torch.from_numpy(synth.synth_tokens(G, 21, ragged=True)).to(dev)
        out["single_process"] = float(solo.contrastive_loss(img_all, tok_all))
        gi_all, gt_all = solo.contrastive_loss_backward()
        grads = [torch.empty(world, b_local, 512, device=dev) for _ in range(2)]
    else:
        gi_all = gt_all = grads = None
    # gather every rank's gradient shard on rank 0 and compare with the single-process gradient
    parts_i = [torch.empty_like(gi) for _ in range(world)] if rank == 0 else None
    parts_t = [torch.empty_like(gt) for _ in range(world)] if rank == 0 else None
    dist.gather(gi, parts_i, dst=0)
    dist.gather(gt, parts_t, dst=0)
    if rank == 0:
        di = torch.cat(parts_i) - gi_all
        dt = torch.cat(parts_t) - gt_all
        out["grad_image_rel"] = float(di.norm() / gi_all.norm())
        out["grad_text_rel"] = float(dt.norm() / gt_all.norm())
    # ---- training step: every rank runs loss_and_backward on its shard (the loss backward reads
    # the peers' row lse in-kernel), the LOCAL parameter gradients are summed with one NCCL all-reduce per tensor (what DDP
    # does), and the result must equal the single-process gradient of the global-batch loss
    if os.environ.get("MSCLIP_PRECISION", "bf16") == "bf16":
        model.enable_training()
        model.zero_grad()
        out["train_loss_p2p"] = float(model.loss_and_backward(img, tok))
        params = model.trainable_parameters()
        uniq = {}
        for k, p in params.items():
            uniq.setdefault(p.data_ptr(), (k, p))
        out["grad_allreduce_collectives"] = model.allreduce_grads()      # bucketed NCCL all-reduce of the library's gradient buffers
        if rank == 0:
            solo.enable_training()
            solo.zero_grad()
            out["train_loss_single"] = float(solo.loss_and_backward(img_all, tok_all))
            ref = solo.trainable_parameters()
            num = den = 0.0
            worst = ("", 0.0)
            for k, p in uniq.values():
                d = float((p.grad - ref[k].grad).double().norm())
                r = float(ref[k].grad.double().norm())
                num += d * d
                den += r * r
                if r > 0 and d / r > worst[1]:
                    worst = (k, d / r)
            out["train_grad_aggregate_rel"] = (num / max(den, 1e-300)) ** 0.5
            out["train_grad_worst"] = list(worst)
    dist.barrier()
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

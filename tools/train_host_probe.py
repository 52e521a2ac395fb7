#!/usr/bin/env python
"""Where does the HOST spend its time in a training step (diagnostic)?  perf_counter around every phase of
loss_and_backward + AdamW for a few steps at batch 4096; a phase that blocks while the GPU idles shows up here."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
from msclip_b200 import synth                  # noqa: E402
from msclip_b200.config import MSCLIPConfig    # noqa: E402
from msclip_b200.model import CLIP             # noqa: E402
from msclip_b200.optim import AdamW            # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = MSCLIPConfig(patch_size=32, layers=12)
sd = synth.synth_state_dict(cfg, seed=0, logit_scale=2.6593)
model = CLIP(cfg, precision="bf16")
model.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=True)
model = model.cuda().eval()
img = torch.randn(B, 3, 224, 224, device="cuda")
tok = torch.from_numpy(synth.synth_tokens(B, 1234, cfg.context_length, cfg.vocab_size)).cuda()
model.enable_training()
opt = AdamW(model, lr=1e-4, weight_decay=0.05, lr_share=1e-4, wd_share=0.2)
for it in range(10):
    t = [time.perf_counter()]
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    opt.zero_grad(); t.append(time.perf_counter())
    loss = model.contrastive_loss(img, tok); t.append(time.perf_counter())
    gi, gt = model.contrastive_loss_backward(); t.append(time.perf_counter())
    model.backward_features(gi, gt); t.append(time.perf_counter())
    fi = model.last_image_features(); model.logit_scale.grad += (gi * fi).sum(); t.append(time.perf_counter())
    # optimiser, split into its two halves
    opt.steps += 1
    n = len(opt.entries)
    P = C.c_void_p
    args = ((P * n)(*[e[1].data_ptr() for e in opt.entries]), (P * n)(*[e[1].grad.data_ptr() for e in opt.entries]),
            (P * n)(*[opt.state[e[0]][0].data_ptr() for e in opt.entries]), (P * n)(*[opt.state[e[0]][1].data_ptr() for e in opt.entries]),
            (C.c_int64 * n)(*[e[1].numel() for e in opt.entries]), (C.c_float * n)(*[e[2] for e in opt.entries]),
            (C.c_float * n)(*[e[3] for e in opt.entries]))
    model._check(model._library().msclip_op_adamw(n, *args, 0.9, 0.999, 1e-8, opt.steps, model._stream()), "adamw"); t.append(time.perf_counter())
    model.refresh_weights([e[0] for e in opt.entries]); t.append(time.perf_counter())
    e1.record()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    names = ["zero_grad", "fwd+loss", "loss_bwd", "backward", "scale_grad", "adamw", "refresh", "final_sync"]
    print(f"step {it}: device {e0.elapsed_time(e1):7.1f} ms | host: " + "  ".join(f"{nm} {1e3 * (b - a):6.1f}" for nm, a, b in zip(names, t, t[1:])), flush=True)

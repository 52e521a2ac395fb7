#!/usr/bin/env python
"""BASELINE.json configs[4]: the zero-shot IN-1K evaluation path at full size on synthetic data -
encode_text over 1000 classes x 80 templates (tools/zero_shot.py:122-134), class-mean + renormalise,
encode_image of a 1024 batch, 100 * I @ W (tools/zero_shot.py:265-266), top-1.

Two ways of building the classifier are timed: the reference's loop (1000 sequential batch-80 calls) and one
batched call over all 80 000 prompts; both must give the same weights.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                               # noqa: E402
import torch                                     # noqa: E402
from msclip_b200 import synth                    # noqa: E402
from msclip_b200.config import MSCLIPConfig      # noqa: E402
from msclip_b200.model import CLIP               # noqa: E402


def main():
    n_cls, n_tpl, n_img = 1000, 80, 1024
    cfg = MSCLIPConfig()
    sd = synth.synth_state_dict(cfg, seed=0)
    model = CLIP(cfg)
    model.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    model = model.cuda().eval()
    toks = torch.from_numpy(synth.synth_tokens(n_cls * n_tpl, 7, ragged=True)).cuda()      # [80000, 77]
    img = torch.randn(n_img, 3, 224, 224, device="cuda")
    model.set_text_trim(False)                                # first the reference's cost model: all 77 positions
    model.encode_text(toks[:80])
    torch.cuda.synchronize()

    t0 = time.perf_counter()
    ws = []
    for c in range(n_cls):                                    # the reference's loop, tools/zero_shot.py:125-131
        e = model.encode_text(toks[c * n_tpl:(c + 1) * n_tpl]).mean(dim=0)
        ws.append(e / e.norm())
    w_loop = torch.stack(ws, dim=0)
    torch.cuda.synchronize()
    t_loop = time.perf_counter() - t0

    t0 = time.perf_counter()
    e = model.encode_text(toks).view(n_cls, n_tpl, -1).mean(dim=1)
    w_batched = e / e.norm(dim=-1, keepdim=True)
    torch.cuda.synchronize()
    t_batched = time.perf_counter() - t0

    # live-prefix path: the causal tower only runs up to the longest EOT position of each call, so the
    # prompts are sorted by length and encoded in length buckets; outputs are bit-identical to the full-context run
    model.set_text_trim(True)
    model.encode_text(toks[:80])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lens = toks.argmax(dim=1)
    order = torch.argsort(lens)
    e_sorted = torch.empty(n_cls * n_tpl, w_batched.shape[1], device="cuda")
    n_buckets = 16
    step = (n_cls * n_tpl + n_buckets - 1) // n_buckets
    for b0 in range(0, n_cls * n_tpl, step):
        sel = order[b0:b0 + step]
        e_sorted[sel] = model.encode_text(toks[sel])
    e = e_sorted.view(n_cls, n_tpl, -1).mean(dim=1)
    w_trim = e / e.norm(dim=-1, keepdim=True)
    torch.cuda.synchronize()
    t_trim = time.perf_counter() - t0
    t0 = time.perf_counter()
    ws = []
    for c in range(n_cls):                                    # the reference's loop again, now on live prefixes
        e = model.encode_text(toks[c * n_tpl:(c + 1) * n_tpl]).mean(dim=0)
        ws.append(e / e.norm())
    w_loop_trim = torch.stack(ws, dim=0)
    torch.cuda.synchronize()
    t_loop_trim = time.perf_counter() - t0

    # fused fast path (msclip_zeroshot_classifier): one library call, sorting / bucketing / class mean inside
    model.zeroshot_classifier(toks.view(n_cls, n_tpl, -1)[:4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    w_fused = model.zeroshot_classifier(toks.view(n_cls, n_tpl, -1)).t().contiguous()
    torch.cuda.synchronize()
    t_fused = time.perf_counter() - t0

    t0 = time.perf_counter()
    feats = model.encode_image(img)
    logits = model.similarity_logits(feats, w_batched, 100.0)
    top1 = logits.argmax(dim=1)
    torch.cuda.synchronize()
    t_img = time.perf_counter() - t0
    t0 = time.perf_counter()
    top5 = model.zeroshot_predict(model.encode_image(img), w_fused.t(), topk=5)
    torch.cuda.synchronize()
    t_img_fused = time.perf_counter() - t0
    out = {"classifier_loop_s": t_loop, "classifier_batched_s": t_batched, "prompts_per_s_loop": n_cls * n_tpl / t_loop,
           "prompts_per_s_batched": n_cls * n_tpl / t_batched, "image_batch_s": t_img, "images_per_s": n_img / t_img,
           "classifier_fused_call_s": t_fused, "prompts_per_s_fused_call": n_cls * n_tpl / t_fused,
           "fused_vs_batched_max_abs_diff": float((w_fused - w_batched).abs().max()),
           "image_batch_fused_predict_s": t_img_fused, "fused_top1_agrees": float((top5[:, 0].long() == top1).float().mean()),
           "classifier_live_prefix_bucketed_s": t_trim, "prompts_per_s_live_prefix_bucketed": n_cls * n_tpl / t_trim,
           "classifier_loop_live_prefix_s": t_loop_trim, "mean_live_length": float(lens.float().mean()) + 1.0,
           "live_prefix_equals_full_context": bool(torch.equal(w_trim, w_batched)) and bool(torch.equal(w_loop_trim, w_loop)),
           "max_abs_diff_loop_vs_batched": float((w_loop - w_batched).abs().max()),
           "logits_shape": list(logits.shape), "top1_agree_loop_vs_batched":
               float((model.similarity_logits(feats, w_loop, 100.0).argmax(dim=1) == top1).float().mean())}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "zeroshot_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

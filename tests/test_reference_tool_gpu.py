"""Row Z of SURVEY.md section 8a, end to end: the reference's UNMODIFIED tools/zero_shot.py (1000 ImageNet classes x 80
templates = 80 000 prompts through encode_text, an ImageFolder through encode_image, `100 * I @ W`, top-1) is run
twice as a subprocess on the same synthetic ImageFolder and the same checkpoint (random init of the REAL reference
model, saved with torch.save(model.state_dict())):

  arm A: the reference's own model (PyTorch eager on the GPU, true fp32: NVIDIA_TF32_OVERRIDE=0),
  arm B: `msclip_b200.dropin` behind `clip_openai_pe_res_v1.get_clip_model` (sitecustomize on PYTHONPATH) - the tool,
         its config system, tokenizer, DataLoader and metric code are byte-identical in both arms.

Checked: the tool built OUR model in arm B, image / class embeddings and the tool's logits agree within the bf16
bounds of tests/test_model_gpu.py, every image whose fp32 top-1 margin exceeds twice the largest logit error gets the
same top-1, and the accuracy the tool PRINTS equals the one recomputed from the recorded tensors in both arms.

Needs the staged reference (tools/stage_reference.py -> baseline/_ref, git-ignored but shipped to the GPU box)."""
import json
import os

import numpy as np
import pytest
import torch

import ref_tool

pytestmark = pytest.mark.gpu
OUT_DIR = os.path.join(ref_tool.REPO, "gpurun_out")


@pytest.mark.parametrize("layers", [12])
def test_unmodified_zero_shot_tool_runs_on_the_dropin(tmp_path, layers):
    root = ref_tool.reference_root()
    if root is None:
        pytest.skip("reference not staged (python tools/stage_reference.py)")
    work = str(tmp_path)
    data = os.path.join(work, "data")
    n_img = ref_tool.make_image_folder(data, n_classes=6, per_class=8, seed=3)
    ckpt = os.path.join(work, "ckpt", "msclips_synth", "model.pth")
    ref_tool.make_checkpoint(ckpt, layers)
    acc_ref, log_ref = ref_tool.run_tool(root, os.path.join(work, "ref"), ckpt, data, layers, dropin=False,
                                         dump=os.path.join(work, "ref.npz"), extra_env={"NVIDIA_TF32_OVERRIDE": "0"})
    acc_our, log_our = ref_tool.run_tool(root, os.path.join(work, "our"), ckpt, data, layers, dropin=True,
                                         dump=os.path.join(work, "our.npz"))
    img_r, w_r, kind_r = ref_tool.load_dump(os.path.join(work, "ref.npz"))
    img_o, w_o, kind_o = ref_tool.load_dump(os.path.join(work, "our.npz"))
    assert kind_r.startswith("models.clip_openai_pe_res_v1") and kind_o == "msclip_b200.model.CLIP", (kind_r, kind_o)
    assert img_r.shape == img_o.shape == (n_img, 512) and w_r.shape == w_o.shape == (1000, 512)

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    logits_r, logits_o = 100.0 * img_r @ w_r.T, 100.0 * img_o @ w_o.T         # tools/zero_shot.py:266
    err = np.abs(logits_o - logits_r)
    top_r, top_o = logits_r.argmax(1), logits_o.argmax(1)
    srt = np.sort(logits_r, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    clear = margin > 2.0 * err.max()
    labels = np.repeat(np.arange(6), 8)                                       # ImageFolder: sorted class dirs
    res = {"image_features": rel(img_o, img_r), "class_embeddings": rel(w_o, w_r), "logits": rel(logits_o, logits_r),
           "logits_max_abs": float(err.max()), "top1_agree": int((top_r == top_o).sum()), "images": int(n_img),
           "clear_margin_images": int(clear.sum()), "clear_agree": int((top_r[clear] == top_o[clear]).sum()),
           "tool_accuracy_reference": acc_ref, "tool_accuracy_dropin": acc_our, "layers": layers,
           "accuracy_from_dump_reference": float((top_r == labels).mean() * 100),
           "accuracy_from_dump_dropin": float((top_o == labels).mean() * 100)}
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, "reference_tool_zero_shot.json"), "w") as f:
        json.dump(res, f, indent=1)
    assert res["image_features"] < 6e-3 and res["class_embeddings"] < 6e-3, res          # FEAT_TOL of test_model_gpu
    assert res["logits_max_abs"] <= 1.5e-3 * 100.0, res                                  # COS_TOL x scale
    assert res["clear_agree"] == res["clear_margin_images"], res
    assert res["top1_agree"] >= 0.9 * n_img, res               # 48 / 48 observed on random-init weights
    assert abs(acc_ref - res["accuracy_from_dump_reference"]) < 1e-3, res
    assert abs(acc_our - res["accuracy_from_dump_dropin"]) < 1e-3, res

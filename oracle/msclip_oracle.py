"""CPU oracle: a functional fp32 restatement of the MS-CLIP-S encode-and-contrast path.

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs; never by the product path (msclip_b200/).

It restates, step by step, the live path of the reference (``M.py`` =
lib/models/clip_openai_pe_res_v1.py of Hxyou/MSCLIP); each function cites the lines it follows.
It is floating-point work, so it is written in torch fp32 (plain ``torch`` ops on a ``dict`` of
tensors — no ``nn.Module``s, batch-major [B, L, D] layout instead of the reference's [L, B, D]).

Parity pinning: tests/golden/*.npz were produced by running the *real* reference
(oracle/ref_shim.py) on msclip_b200.synth weights/inputs (oracle/make_golden.py);
tests/test_oracle.py checks this file against them.  The reference itself ships no tests or golden
vectors, and the loss is not in the reference at all (M.py:3155 returns logits): the
loss oracle is the standard symmetric cross-entropy over the reference's logits, as the north star
specifies.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

LATERAL_LAYERS = (2, 4, 6, 8, 10)       # b32-yfcc-msclips.yaml:18


# ----------------------------------------------------------------------------- primitives
def layer_norm(x, w, b, eps: float = 1e-12):
    """TF-style LayerNorm, eps inside the sqrt, fp32 statistics (M.py:204-219)."""
    x = x.float()
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return w * ((x - u) / torch.sqrt(s + eps)) + b


def quick_gelu(x):
    """x * sigmoid(1.702 x) (M.py:222-224)."""
    return x * torch.sigmoid(1.702 * x)


def batch_norm_eval(x, sd, prefix: str, eps: float):
    """BatchNorm2d in eval mode: running statistics (M.py BN modules; eps 1e-5 default, 1e-6 in
    ConvResBlock M.py:1825-1840)."""
    m, v = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    g, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    scale = g / torch.sqrt(v + eps)
    return x * scale[None, :, None, None] + (b - m * scale)[None, :, None, None]


def causal_mask(n: int, device=None):
    """Additive mask, -inf strictly above the diagonal (M.py:2965-2971)."""
    return torch.full((n, n), float("-inf"), device=device).triu_(1)


def attention(x, w_in, b_in, w_out, b_out, heads: int, mask: Optional[torch.Tensor]):
    """Self-attention of Attention_CUST.forward, live lines M.py:612, 707-738, 747.

    x: [B, L, D].  q is scaled by head_dim^-0.5 *after* the bias add (M.py:612 then 707).
    Head h owns channels [64h, 64h+64) (view(L, B*H, 64), M.py:709-711).
    """
    B, L, D = x.shape
    hd = D // heads
    qkv = F.linear(x, w_in, b_in)
    q, k, v = qkv.split(D, dim=-1)
    q = q * (float(hd) ** -0.5)
    q = q.view(B, L, heads, hd).transpose(1, 2)
    k = k.view(B, L, heads, hd).transpose(1, 2)
    v = v.view(B, L, heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2)                       # [B, H, L, L]   (bmm, M.py:716)
    if mask is not None:
        s = s + mask                                  # M.py:725
    p = torch.softmax(s, dim=-1)                      # M.py:728 (dropout p=0, M.py:734)
    o = (p @ v).transpose(1, 2).reshape(B, L, D)      # M.py:736-738
    return F.linear(o, w_out, b_out)                  # M.py:747


def residual_block(x, sd: Dict[str, torch.Tensor], prefix: str, heads: int, mask):
    """Pre-LN block: x += attn(ln_1(x)); x += mlp(ln_2(x)) (M.py:1027-1028; mlp M.py:794-798)."""
    a = attention(layer_norm(x, sd[prefix + ".ln_1.weight"], sd[prefix + ".ln_1.bias"]),
                  sd[prefix + ".attn.in_proj_weight"], sd[prefix + ".attn.in_proj_bias"],
                  sd[prefix + ".attn.out_proj.weight"], sd[prefix + ".attn.out_proj.bias"], heads, mask)
    x = x + a
    h = layer_norm(x, sd[prefix + ".ln_2.weight"], sd[prefix + ".ln_2.bias"])
    h = F.linear(h, sd[prefix + ".mlp.c_fc.weight"], sd[prefix + ".mlp.c_fc.bias"])
    h = quick_gelu(h)
    h = F.linear(h, sd[prefix + ".mlp.c_proj.weight"], sd[prefix + ".mlp.c_proj.bias"])
    return x + h


# ----------------------------------------------------------------------------- vision pieces
def early_conv_stem(img, sd, prefix: str, strides):
    """EarlyconvRes.forward (M.py:1993-2000) with ResBasicBlock_v0 (M.py:1920-1936)."""
    x = F.conv2d(img, sd[prefix + "conv1.weight"], stride=2, padding=1)
    x = torch.relu(batch_norm_eval(x, sd, prefix + "bn1", 1e-5))
    for i, s in enumerate(strides):
        p = prefix + f"resnet_stage.conv_{i}."
        main = batch_norm_eval(F.conv2d(x, sd[p + "conv1.weight"], stride=s, padding=1), sd, p + "bn1", 1e-5)
        skip = batch_norm_eval(F.conv2d(x, sd[p + "downsample.0.weight"], stride=s), sd, p + "downsample.1", 1e-5)
        x = torch.relu(main + skip)
    return F.conv2d(x, sd[prefix + "last_conv.weight"])


def branch_stage(x, sd, prefix: str, j: int, stride: int):
    """Parallel-branch stage j (M.py:2436/2442): j=0 conv-bn-relu (M.py:2260-2273); j>=1 one
    ConvResBlock with projection shortcut (M.py:1842-1861), BN eps 1e-6."""
    if j == 0:
        x = F.conv2d(x, sd[prefix + "0.conv.weight"], stride=stride, padding=1)
        return torch.relu(batch_norm_eval(x, sd, prefix + "0.bn", 1e-5))
    p = prefix + f"{j}.resnet_stage.conv_0."
    y = torch.relu(batch_norm_eval(F.conv2d(x, sd[p + "conv1.weight"]), sd, p + "bn1", 1e-6))
    y = torch.relu(batch_norm_eval(F.conv2d(y, sd[p + "conv2.weight"], stride=stride, padding=1), sd, p + "bn2", 1e-6))
    y = batch_norm_eval(F.conv2d(y, sd[p + "conv3.weight"]), sd, p + "bn3", 1e-6)
    r = batch_norm_eval(F.conv2d(x, sd[p + "residual_conv.weight"], stride=stride), sd, p + "residual_bn", 1e-6)
    return torch.relu(y + r)


def lateral_adapter(top, x, sd, prefix: str, k: int, grid: int):
    """Lateral_Adapter.forward, top->bottom only (M.py:1752-1778).

    top: branch feature [B, C, H, W]; x: tokens [B, 1+grid^2, D].  The adapter *replaces* x (there
    is no skip from x) and, with PRALLEL_T2B_USECLS, the CLS row becomes ln_adapt(cls + cls).
    """
    B, L, D = x.shape
    C = top.shape[1]
    t = F.conv2d(top, sd[prefix + "top2bottom_dw_conv.conv.weight"], stride=k, groups=C)
    t = batch_norm_eval(t, sd, prefix + "top2bottom_dw_conv.bn", 1e-5)
    t = F.conv2d(t, sd[prefix + "top2bottom_pw_conv.conv.weight"])            # [B, D, g, g]
    t = t.flatten(2).transpose(1, 2)                                           # [B, g*g, D]
    cls, tok = x[:, :1], x[:, 1:]
    gmap = tok.transpose(1, 2).reshape(B, D, grid, grid)
    b = F.conv2d(gmap, sd[prefix + "bottom_dw_conv.conv.weight"], padding=1, groups=D)
    b = batch_norm_eval(b, sd, prefix + "bottom_dw_conv.bn", 1e-5).flatten(2).transpose(1, 2)
    y = torch.cat([cls + cls, b + t], dim=1)
    return layer_norm(y, sd[prefix + "ln_adapt.weight"], sd[prefix + "ln_adapt.bias"])


def vision_tokens(img, sd, cfg, taps: Optional[dict] = None):
    """Vision tower up to (and including) the last block: Transformer.forward with first_conv
    (M.py:2388-2459).  Returns tokens [B, L, D]."""
    v = "visual."
    x = early_conv_stem(img, sd, v + "transformer.resblocks.0.", cfg.early_strides)
    B = x.shape[0]
    x = x.flatten(2).transpose(1, 2)                                           # M.py:2418-2419
    cls = sd[v + "class_embedding"].expand(B, 1, -1)
    x = torch.cat([cls, x], dim=1) + sd[v + "positional_embedding"]          # M.py:2420-2423
    x = layer_norm(x, sd[v + "ln_pre.weight"], sd[v + "ln_pre.bias"])          # M.py:2424
    if taps is not None:
        taps["v_embed"] = x
    par = img
    for idx in range(1, cfg.layers):
        if idx in LATERAL_LAYERS:
            j = LATERAL_LAYERS.index(idx)
            par = branch_stage(par, sd, v + "transformer.parallel_branch_v.", j, cfg.parallel_strides[j])
            x = lateral_adapter(par, x, sd, v + f"transformer.parallel_lateral_adapter.{j}.",
                                cfg.t2b_kernels[j], cfg.grid)
            if taps is not None:
                taps[f"v_adapter{j}"] = x
        x = residual_block(x, sd, v + f"transformer.resblocks.{idx}", cfg.heads, None)
        if taps is not None:
            taps[f"v_block{idx}"] = x
    return x


def encode_image(img, sd, cfg, norm: bool = True, taps: Optional[dict] = None):
    """CLIP.encode_image (M.py:2979-2985) = VisualTransformer.forward (M.py:2621-2697)."""
    v = "visual."
    x = vision_tokens(img.float(), sd, cfg, taps)
    x = layer_norm(x[:, 0], sd[v + "ln_post.weight"], sd[v + "ln_post.bias"])   # M.py:2685-2687
    x = x @ sd[v + "proj"]                                                     # M.py:2690
    if norm:
        x = x / x.norm(dim=-1, keepdim=True)                                   # M.py:2982-2983
    return x


def encode_text(tok, sd, cfg, norm: bool = True, taps: Optional[dict] = None):
    """CLIP.encode_text (M.py:3043-3079)."""
    x = sd["token_embedding.weight"][tok] + sd["positional_embedding"]        # M.py:3047-3048
    mask = causal_mask(tok.shape[1], x.device)
    for idx in range(cfg.layers):
        x = residual_block(x, sd, f"transformer.resblocks.{idx}", cfg.heads, mask)
        if taps is not None:
            taps[f"t_block{idx}"] = x
    eot = tok.argmax(dim=-1)                                                   # M.py:3059
    x = x[torch.arange(x.shape[0], device=x.device), eot]
    x = layer_norm(x, sd["ln_final.weight"], sd["ln_final.bias"])              # M.py:3072
    x = x @ sd["text_projection"]                                              # M.py:3074
    if norm:
        x = x / x.norm(dim=-1, keepdim=True)                                   # M.py:3076-3077
    return x


# ----------------------------------------------------------------------------- contrast
def similarity_logits(f_img, f_txt, logit_scale):
    """exp(logit_scale) * I @ T^T (M.py:3136, 3141/3146); rows = images."""
    t = logit_scale.exp() if torch.is_tensor(logit_scale) else math.exp(logit_scale)
    return t * f_img @ f_txt.t()


def contrastive_loss(logits):
    """Symmetric cross-entropy over the logits (not in the reference: M.py:3155 returns logits;
    spec = north star / SURVEY.md §8(a) row L)."""
    target = torch.arange(logits.shape[0], device=logits.device)
    return 0.5 * (F.cross_entropy(logits, target) + F.cross_entropy(logits.t(), target))


def forward(img, tok, sd, cfg):
    """CLIP.forward (M.py:3126-3155), single process (gather = identity)."""
    fi, ft = encode_image(img, sd, cfg), encode_text(tok, sd, cfg)
    return similarity_logits(fi, ft, sd["logit_scale"])


def gather_rank_order(shards):
    """gather_tensors (lib/utils/comm.py:140-154): rank-ordered concat along dim 0."""
    return torch.cat(list(shards), dim=0)


def zeroshot_classifier(class_tokens, sd, cfg):
    """tools/zero_shot.py:122-134: per class mean of normalised prompt embeddings, renormalised;
    class_tokens: [n_classes, n_templates, ctx] -> weights [E, n_classes]."""
    ws = []
    for toks in class_tokens:
        e = encode_text(toks, sd, cfg)
        e = e.mean(dim=0)
        ws.append(e / e.norm())
    return torch.stack(ws, dim=1)


def zeroshot_logits(img, weights, sd, cfg):
    """tools/zero_shot.py:265-266: 100 * encode_image(x) @ W."""
    return 100.0 * encode_image(img, sd, cfg) @ weights


def to_torch(sd_np, device="cpu"):
    """numpy state dict (msclip_b200.synth) -> torch fp32 tensors, aliases preserved."""
    cache, out = {}, {}
    for k, a in sd_np.items():
        if id(a) not in cache:
            cache[id(a)] = torch.as_tensor(a).to(device)
        out[k] = cache[id(a)]
    return out

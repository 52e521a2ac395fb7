#!/usr/bin/env python
"""Throughput of the GPU input transform against the tool's torchvision transform on the host cores.

ImageNet-like sizes (500 x 375 and friends), batch of `--n` images: msclip_preprocess_images from pinned host bytes (H2D copy of
the decoded pixels inside the timed region) and from device-resident bytes; comparator = tools/zero_shot.py:202-207's
transforms.Compose on PIL images, one image at a time on one core (what a DataLoader worker does).
Writes gpurun_out/preprocess_bench.json.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                              # noqa: E402
import torch                                    # noqa: E402
from msclip_b200.config import MSCLIPConfig     # noqa: E402
from msclip_b200.model import CLIP              # noqa: E402


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)      # lib/config/default.py:84-85
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    r = np.random.RandomState(0)
    sizes = [(375, 500), (500, 375), (333, 500), (500, 500), (224, 224), (768, 1024)]
    imgs = [r.randint(0, 256, size=sizes[i % len(sizes)] + (3,)).astype(np.uint8) for i in range(args.n)]
    src_bytes = sum(im.size for im in imgs)
    model = CLIP(MSCLIPConfig(patch_size=32, layers=2)).cuda().eval()
    out = {"n": args.n, "source_MB": src_bytes / 1e6}
    for name, on_dev in (("host_pinned_source", False), ("device_source", True)):
        model.preprocess(imgs, source_on_device=on_dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            model.preprocess(imgs, source_on_device=on_dev)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.reps
        out[name] = {"ms": dt * 1e3, "images_per_s": args.n / dt, "note": "wall clock incl. host-side packing of the list into one buffer"}
    # kernels only (device-resident packed bytes, CUDA events around the C-ABI call)
    import ctypes as C
    from msclip_b200 import _lib
    L = _lib.lib("bf16")
    offs, total = [], 0
    for im in imgs:
        offs.append(total)
        total += im.size
    packed = torch.from_numpy(np.concatenate([im.reshape(-1) for im in imgs])).cuda()
    res = torch.empty(args.n, 3, 224, 224, device="cuda")
    n = args.n
    call = lambda: _lib.check(L.msclip_preprocess_images(
        model._ensure_handle(), C.c_void_p(packed.data_ptr()), (C.c_int64 * n)(*offs), (C.c_int * n)(*[im.shape[0] for im in imgs]),
        (C.c_int * n)(*[im.shape[1] for im in imgs]), n, 224, (C.c_float * 3)(*CLIP_MEAN), (C.c_float * 3)(*CLIP_STD),
        C.c_void_p(res.data_ptr()), _lib.F32, None, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    out["c_abi_device_source"] = {"ms": ms, "images_per_s": n / ms * 1e3, "source_GBps": src_bytes / ms / 1e6,
                                  "note": "includes the host-side coefficient tables (double arithmetic) and their upload"}
    from PIL import Image
    from torchvision import transforms
    t = transforms.Compose([transforms.Resize(224, interpolation=Image.BICUBIC), transforms.CenterCrop((224, 224)), transforms.ToTensor(),
                            transforms.Normalize(mean=CLIP_MEAN, std=CLIP_STD)])
    pil = [Image.fromarray(im) for im in imgs[:64]]
    t0 = time.perf_counter()
    for im in pil:
        t(im)
    dt = (time.perf_counter() - t0) / len(pil)
    out["torchvision_one_core"] = {"ms_per_image": dt * 1e3, "images_per_s": 1 / dt}
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "preprocess_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

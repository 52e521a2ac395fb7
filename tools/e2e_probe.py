"""Diagnostic: what does a concurrent pinned H2D copy cost the library's kernels?"""
import ctypes as C, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msclip_b200 import _lib, synth
from msclip_b200.config import MSCLIPConfig
from msclip_b200.model import CLIP
cfg = MSCLIPConfig(layers=12)
sd = synth.synth_state_dict(cfg, seed=0)
m = CLIP(cfg); m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}); m = m.cuda().eval(); m._sync_weights()
B = 4096
img = torch.randn(B, 3, 224, 224).pin_memory()
imgd = img.cuda()
scratch = torch.empty_like(imgd)
tok = torch.from_numpy(synth.synth_tokens(B, 1)).cuda()
side = torch.cuda.Stream()


def timed(fn, copy):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if copy:
        with torch.cuda.stream(side):
            scratch.copy_(img, non_blocking=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for name, fn in (("encode_text", lambda: m.encode_text(tok)), ("encode_image", lambda: m.encode_image(imgd)),
                 ("contrastive_loss(device inputs)", lambda: m.contrastive_loss(imgd, tok))):
    fn()
    a = min(timed(fn, False) for _ in range(3))
    b = min(timed(fn, True) for _ in range(3))
    print(f"{name}: alone {a:.1f} ms, with one concurrent 2.47 GB H2D copy (torch side stream) {b:.1f} ms")

L = _lib.lib(); h = m._handle; sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
tokh = tok.cpu().pin_memory(); out = torch.zeros(3).pin_memory()
def stage():
    _lib.check(L.msclip_stage_images(h, C.c_void_p(img.data_ptr()), 0, B, sp))
def fwd(ptr):
    t0 = time.perf_counter()
    _lib.check(L.msclip_forward_loss(h, C.c_void_p(ptr), 0, C.c_void_p(tokh.data_ptr()), B, C.c_void_p(out.data_ptr()), C.c_void_p(out.data_ptr() + 8), sp))
    return (time.perf_counter() - t0) * 1e3
print("loop: stage(next) then forward_loss(host images, pinned host tokens), no sync in between")
torch.cuda.synchronize(); stage()
for i in range(6):
    stage(); print(f"  step {i}: forward_loss {fwd(img.data_ptr()):.1f} ms")
torch.cuda.synchronize()
for i in range(3):
    print(f"  device images: forward_loss {fwd(imgd.data_ptr()):.1f} ms")

"""Input transform (tools/zero_shot.py:202-207): byte work, so every comparison is BIT-EXACT.

CPU (not gpu): the oracle (oracle/preprocess_oracle.py) against the committed fixture and against the live torchvision +
Pillow stack on random sizes.  GPU: msclip_preprocess_images (through the C ABI) against the oracle, the fixture and
torchvision, for mixed-size batches, host and device sources, all output dtypes.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as P

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess_cases.npz")


def torchvision_transform(img, size=224):
    from PIL import Image
    from torchvision import transforms
    t_u8 = transforms.Compose([transforms.Resize(size, interpolation=Image.BICUBIC), transforms.CenterCrop((size, size))])
    pil = Image.fromarray(img)
    u8 = np.asarray(t_u8(pil))
    f = transforms.Normalize(mean=P.CLIP_MEAN, std=P.CLIP_STD)(transforms.ToTensor()(t_u8(pil))).numpy()
    return u8, f


def test_oracle_matches_committed_fixture():
    z = np.load(GOLDEN)
    for i, (h, w) in enumerate(P.FIXTURE_SIZES):
        img = P.fixture_image(h, w, i)
        assert np.array_equal(P.resize_center_crop_u8(img, 224), z[f"u8_{h}x{w}"]), (h, w)
        assert np.array_equal(P.transform(img, 224, P.CLIP_MEAN, P.CLIP_STD)[:, ::7, ::5], z[f"f32_{h}x{w}"]), (h, w)


@pytest.mark.parametrize("h,w,size", [(224, 224, 224), (240, 320, 224), (333, 250, 224), (48, 48, 224), (37, 91, 32), (600, 601, 224),
                                      (225, 224, 224), (1, 5, 8)])
def test_oracle_matches_live_torchvision(h, w, size):
    img = np.random.RandomState(h * 1000 + w).randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    u8, f = torchvision_transform(img, size)
    assert np.array_equal(P.resize_center_crop_u8(img, size), u8)
    assert np.array_equal(P.transform(img, size, P.CLIP_MEAN, P.CLIP_STD), f)


# ------------------------------------------------------------------------------------------------------- GPU
def _gpu_preprocess(model, images, size=224, dtype=torch.float32, on_device=False, want_u8=True):
    return model.preprocess(images, size=size, dtype=dtype, source_on_device=on_device, return_u8=want_u8)


@pytest.fixture(scope="module")
def model():
    from msclip_b200.config import MSCLIPConfig
    from msclip_b200.model import CLIP
    return CLIP(MSCLIPConfig(patch_size=32, layers=2)).cuda().eval()


@pytest.mark.gpu
def test_gpu_transform_is_bit_exact_on_the_fixture(model):
    z = np.load(GOLDEN)
    imgs = [P.fixture_image(h, w, i) for i, (h, w) in enumerate(P.FIXTURE_SIZES)]
    out, u8 = _gpu_preprocess(model, imgs)
    for i, (h, w) in enumerate(P.FIXTURE_SIZES):
        assert np.array_equal(u8[i].cpu().numpy(), z[f"u8_{h}x{w}"]), (h, w)
        assert np.array_equal(out[i].cpu().numpy()[:, ::7, ::5], z[f"f32_{h}x{w}"]), (h, w)


@pytest.mark.gpu
@pytest.mark.parametrize("on_device", [False, True])
def test_gpu_transform_matches_torchvision_on_mixed_batches(model, on_device):
    r = np.random.RandomState(7)
    sizes = [(224, 224), (256, 341), (500, 333), (75, 75), (231, 640), (640, 231), (225, 224), (1200, 900), (224, 225), (33, 47)]
    imgs = [r.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w in sizes]
    out, u8 = _gpu_preprocess(model, imgs, on_device=on_device)
    for i, img in enumerate(imgs):
        ref_u8, ref_f = torchvision_transform(img)
        assert np.array_equal(u8[i].cpu().numpy(), ref_u8), sizes[i]
        assert np.array_equal(out[i].cpu().numpy(), ref_f), sizes[i]
        assert np.array_equal(out[i].cpu().numpy(), P.transform(img, 224, P.CLIP_MEAN, P.CLIP_STD)), sizes[i]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_gpu_transform_16bit_outputs_are_the_rounded_fp32(model, dtype):
    r = np.random.RandomState(8)
    imgs = [r.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w in [(300, 200), (224, 400)]]
    f32, _ = _gpu_preprocess(model, imgs)
    f16, _ = _gpu_preprocess(model, imgs, dtype=dtype, want_u8=False)
    assert f16.dtype == dtype and torch.equal(f16, f32.to(dtype))


@pytest.mark.gpu
def test_preprocessed_batch_feeds_encode_image(model):
    """The tool's loop: transform -> encode_image (tools/zero_shot.py:262-266) entirely on the device."""
    r = np.random.RandomState(9)
    imgs = [r.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w in [(300, 200), (224, 400), (250, 250)]]
    x, _ = _gpu_preprocess(model, imgs, want_u8=False)
    ref = torch.stack([torch.from_numpy(torchvision_transform(im)[1]) for im in imgs]).cuda()
    assert torch.equal(model.encode_image(x), model.encode_image(ref))

"""Cosine schedule of experiments/model/b32.yaml:37-46 (timm 'cosine': warm-up 5 epochs from 1e-6, min_lr 1e-5, cool-down 10) on
the CPU: the rule is restated from timm's CosineLRScheduler (not installed here), so the test pins the formula's fixed points."""
import math

from msclip_b200.optim import CosineSchedule


class _FakeOpt:
    def __init__(self):
        self.entries = [["a", None, 1e-4, 0.05], ["visual.transformer.resblocks.1.mlp.c_fc.weight", None, 2e-4, 0.2]]


def test_cosine_schedule_fixed_points_and_groups():
    opt = _FakeOpt()
    s = CosineSchedule(opt, epochs=32, warmup_epochs=5, warmup_lr=1e-6, min_lr=1e-5, cooldown_epochs=10)
    assert s.total_epochs() == 42
    s.step(0)
    assert [e[2] for e in opt.entries] == [1e-6, 1e-6]
    s.step(2.5)                                                   # half way through the warm-up
    assert math.isclose(opt.entries[0][2], 1e-6 + 0.5 * (1e-4 - 1e-6)) and math.isclose(opt.entries[1][2], 1e-6 + 0.5 * (2e-4 - 1e-6))
    s.step(5)                                                     # first cosine epoch: timm counts t from epoch 0
    want = 1e-5 + 0.5 * (1e-4 - 1e-5) * (1 + math.cos(math.pi * 5 / 32))
    assert math.isclose(opt.entries[0][2], want)
    s.step(16)                                                    # mid cycle: half way between base and min
    assert math.isclose(opt.entries[0][2], 1e-5 + 0.5 * (1e-4 - 1e-5)) and math.isclose(opt.entries[1][2], 1e-5 + 0.5 * (2e-4 - 1e-5))
    for ep in (32, 35, 41.9):                                     # cool-down at min_lr
        s.step(ep)
        assert [e[2] for e in opt.entries] == [1e-5, 1e-5]
    lrs = []
    for ep in range(5, 33):
        s.step(ep)
        lrs.append(opt.entries[0][2])
    assert all(a > b for a, b in zip(lrs, lrs[1:]))               # monotone decay over the cycle

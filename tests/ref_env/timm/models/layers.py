import torch


class DropPath(torch.nn.Module):
    """Stochastic depth; the MS-CLIP-S configs run it at p = 0 (identity)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob or 0.0)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


trunc_normal_ = torch.nn.init.trunc_normal_

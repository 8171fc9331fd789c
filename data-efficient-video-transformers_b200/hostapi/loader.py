"""Input-side feature path on the GPU (src/dataloaders/MMX_Temporal_dl.py:167-181).

The reference's Dataset pads every expert vector to 2048 on the CPU and, in the train state, zeroes it with probability
0.3 and adds N(0, 0.1) noise with probability 0.3, one vector at a time in DataLoader workers.  ``FeatureAugment`` does
the same per (clip, scene) vector in one kernel launch per expert tensor after the host->device copy of the RAW
features, and writes the activation dtype the embed prologue wants (so the separate fp32 -> bf16 cast disappears).
``pad_to=None`` keeps an expert's own width - the input projection then never multiplies zero columns."""
import torch

from .. import ops
from ..functions import next_seed


class FeatureAugment(torch.nn.Module):
    def __init__(self, p_drop=0.3, p_noise=0.3, noise_std=0.1 ** 0.5, pad_to=None, out_dtype=torch.bfloat16):
        super().__init__()
        self.p_drop, self.p_noise, self.noise_std, self.pad_to, self.out_dtype = p_drop, p_noise, noise_std, pad_to, out_dtype

    def forward(self, feats, seed=None):
        """feats: a [B, T, D_e] fp32 CUDA tensor or a list of them (one per expert)."""
        if isinstance(feats, (list, tuple)):
            return [self.forward(f, None if seed is None else seed + i) for i, f in enumerate(feats)]
        train = self.training                                      # the loader transforms only in its "train" state (:172)
        return ops.feature_augment(feats, self.pad_to, p_drop=self.p_drop if train else 0.0,
                                   p_noise=self.p_noise if train else 0.0, noise_std=self.noise_std,
                                   seed=next_seed() if seed is None else seed, out_dtype=self.out_dtype)

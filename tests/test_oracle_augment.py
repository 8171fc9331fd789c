"""CPU: the augmentation oracle (oracle/augment.py).  ``apply`` against the reference's own statements executed vector
by vector (src/dataloaders/MMX_Temporal_dl.py:167-169,176-181 with the random draws replaced by given decisions), and
the counter-based stream's distribution."""
import numpy as np
import torch
import torch.nn as nn

from oracle import augment


def test_apply_is_the_reference_transform():
    rng = np.random.default_rng(0)
    rows, d_in, d_out = 64, 128, 2048
    x = rng.standard_normal((rows, d_in)).astype(np.float32)
    drop, noisy = rng.random(rows) < 0.3, rng.random(rows) < 0.3
    z = rng.standard_normal((rows, d_out)).astype(np.float32) * np.float32(0.1 ** 0.5)
    got = augment.apply(x, d_out, drop, noisy, z)
    for r in range(rows):
        t = torch.from_numpy(x[r:r + 1])
        if t.shape[-1] != 2048:
            t = nn.ConstantPad1d((0, 2048 - t.shape[-1]), 0)(t)            # :167-169
        if drop[r]:
            t = torch.zeros((1, 2048))                                      # :177-178
        if noisy[r]:
            t = t + torch.from_numpy(z[r:r + 1])                            # :179-180 (noise given)
        assert np.array_equal(got[r:r + 1], t.numpy())


def test_counter_stream_distribution_and_determinism():
    seed = 0xDEADBEEF12345678
    d1, n1 = augment.decisions(seed, 200000, 0.3, 0.3)
    d2, n2 = augment.decisions(seed, 200000, 0.3, 0.3)
    assert np.array_equal(d1, d2) and np.array_equal(n1, n2)
    assert abs(d1.mean() - 0.3) < 0.005 and abs(n1.mean() - 0.3) < 0.005 and abs((d1 & n1).mean() - 0.09) < 0.004
    d3, _ = augment.decisions(seed + 1, 200000, 0.3, 0.3)
    assert abs((d1 == d3).mean() - (0.7 * 0.7 + 0.3 * 0.3)) < 0.005          # a different seed is an independent stream
    z = augment.noise(seed, 4000, 64, 0.1 ** 0.5)
    assert abs(z.var() - 0.1) < 0.002 and abs(z.mean()) < 0.002
    assert abs(np.mean(z ** 4) / z.var() ** 2 - 3.0) < 0.1                    # Gaussian kurtosis

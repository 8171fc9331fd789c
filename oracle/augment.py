"""TEST INFRASTRUCTURE ONLY (CPU oracle; never imported by the product path).

Loader-side feature path of the reference, src/dataloaders/MMX_Temporal_dl.py:
  :167-169  ``if t.shape[-1] != 2048: t = nn.ConstantPad1d((0, 2048 - t.shape[-1]), 0)(t)``   zero-pad to 2048
  :176-181  ``add_transforms``: ``if random.random() < 0.3: x = zeros((1, 2048))`` then
            ``if random.random() < 0.3: x = x + (0.1 ** 0.5) * torch.randn(1, 2048)``          (train state only, :172-173)
The reference draws its decisions from Python's / torch's global CPU generators, which no GPU kernel can replay; what
is pinned is (a) the transform given the decisions (``apply``, a literal restatement of the lines above), and (b) the
counter-based random stream the CUDA kernel uses (``decisions`` / ``noise``: the same 5-round multiply-xor hash as
``tvt_common.cuh::dropout_words`` restated in numpy), so the kernel is compared bit for bit on every drop / noise
decision and to float rounding on the noise values.  The distribution (rates 0.3 / 0.3, variance 0.1, independence)
is checked statistically in the tests."""
import numpy as np

ROUNDS, MUL, WEYL, NOISE_STREAM = 5, 0xD256D193, 0x9E3779B9, 0x6E6F6973
M32 = 0xFFFFFFFF


def hash_words(seed, counter, stream=0):
    """(w_lo, w_hi) uint32 arrays for uint64 counters: tvt_common.cuh::dropout_words with round keys seed_lo + r * WEYL
    and starting words R = counter_lo, L = counter_hi ^ seed_hi ^ stream."""
    counter = np.asarray(counter, dtype=np.uint64)
    R = (counter & np.uint64(M32)).astype(np.uint64)
    L = ((counter >> np.uint64(32)) ^ np.uint64(((seed >> 32) ^ stream) & M32)).astype(np.uint64)
    for r in range(ROUNDS):
        m = R * np.uint64(MUL)                                   # < 2^64: both factors are < 2^32
        R = ((m >> np.uint64(32)) ^ np.uint64((seed + r * WEYL) & M32) ^ L) & np.uint64(M32)
        L = m & np.uint64(M32)
    return R.astype(np.uint32), L.astype(np.uint32)


def u01(w):
    """Open-interval uniform from the top 24 bits, in float32 arithmetic exactly as the kernel does it."""
    return (np.float32(1) * (w >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def decisions(seed, rows, p_drop, p_noise):
    w0, w1 = hash_words(seed, np.arange(rows, dtype=np.uint64))
    return u01(w0) < np.float32(p_drop), u01(w1) < np.float32(p_noise)


def noise(seed, rows, d_out, std):
    """[rows, d_out] float32 Gaussian noise: one hash per column pair, Box-Muller (cos -> even column, sin -> odd)."""
    ctr = np.arange(rows * (d_out // 2), dtype=np.uint64)
    a, b = hash_words(seed, ctr, NOISE_STREAM)
    rad = np.sqrt(np.float32(-2.0) * np.log(u01(a))) * np.float32(std)
    ang = np.float32(6.283185307179586) * u01(b)
    z = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=-1).astype(np.float32)
    return z.reshape(rows, d_out)


def apply(x, d_out, drop, add_noise, z):
    """MMX_Temporal_dl.py:167-169,176-181 given the decisions: pad, zero the dropped vectors, add noise to the chosen ones."""
    rows, d_in = x.shape
    out = np.zeros((rows, d_out), dtype=np.float32)
    out[:, :d_in] = x
    out[drop] = 0.0
    out[add_noise] += z[add_noise]
    return out


def feature_augment(x, d_out, p_drop, p_noise, std, seed):
    x = np.asarray(x, dtype=np.float32)
    drop, add_noise = decisions(seed, x.shape[0], p_drop, p_noise)
    return apply(x, d_out, drop, add_noise, noise(seed, x.shape[0], d_out, std)), drop, add_noise

"""Counter-based prenet-dropout keep mask (numpy) -- TEST INFRASTRUCTURE.

The reference prenet applies F.dropout at inference with torch's global RNG
(nets/modules/decoder_sa.py:156-157), so its output is not reproducible across
implementations. The B200 path defines the keep decision as a pure function

    words = Philox4x32-10(key = (seed & 0xffffffff, seed >> 32),
                          counter = (unit >> 3, step | layer << 24, phoneme, utt))
    j = unit & 7;  lane16 = (words[j >> 1] >> 16) if (j & 1) else (words[j >> 1] & 0xffff)
    keep(seed, utt, phoneme, step, layer, unit) = lane16 >= thresh16(p)

with thresh16(p) = min(floor(float32(p) * 65536 + 0.5), 65535): one Philox call decides 8 units
(16-bit resolution of p; exact for the reference's p = 0.5). This file restates that
definition on the CPU (the CUDA copy is fcl_taco2_b200/csrc/philox.cuh); the
oracle injects it into the reference through oracle.ref_loader.prenet_dropout.
Philox4x32-10 is Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"
(SC'11); the known-answer vectors in tests/test_oracle_philox.py are the
Random123 ones.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All args broadcastable uint32 arrays/ints -> 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)])
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for r in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def threshold16(p: float) -> int:
    """p is taken as float32 (what the CUDA side receives): min(floor(float32(p) * 65536 + 0.5), 65535)."""
    return min(int(np.float32(np.float32(p) * np.float32(65536.0) + np.float32(0.5))), 65535)


def keep_mask(seed: int, utt, phoneme, step: int, layer: int, n_units: int, p: float) -> np.ndarray:
    """-> bool (rows, n_units). `utt`, `phoneme`: int arrays (rows,)."""
    assert n_units % 8 == 0
    utt = np.asarray(utt, dtype=np.uint64).reshape(-1, 1)
    ph = np.asarray(phoneme, dtype=np.uint64).reshape(-1, 1)
    octs = np.arange(n_units // 8, dtype=np.uint64).reshape(1, -1)
    c1 = np.uint64((step & 0xFFFFFF) | (layer << 24))
    w = philox4x32_10(octs, c1, ph, utt, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(w, axis=-1)                                               # (rows, octs, 4)
    lanes = np.stack([words & np.uint32(0xFFFF), words >> np.uint32(16)], axis=-1)   # (rows, octs, 4, 2): unit j -> [j>>1][j&1]
    lanes = lanes.reshape(utt.shape[0], n_units)
    return lanes >= np.uint32(threshold16(p))

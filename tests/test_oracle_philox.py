"""Known-answer tests pinning the counter-based dropout RNG (Random123 kat vectors)."""
import numpy as np

from oracle import philox


def _k(c, k):
    return tuple(int(v) for v in philox.philox4x32_10(*c, *k))


def test_philox_kat():
    assert _k((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    f = 0xFFFFFFFF
    assert _k((f, f, f, f), (f, f)) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert _k((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_keep_mask_rate_and_invariance():
    m = philox.keep_mask(7, np.zeros(64, int), np.arange(64), 3, 1, 256, 0.5)
    assert m.shape == (64, 256) and 0.47 < m.mean() < 0.53
    # row r of a batch equals the same (utt, phoneme) evaluated alone: batching/sorting invariant
    one = philox.keep_mask(7, [0], [17], 3, 1, 256, 0.5)
    assert (one[0] == m[17]).all()
    assert philox.threshold16(0.5) == 1 << 15 and philox.threshold16(1.0) == 0xFFFF

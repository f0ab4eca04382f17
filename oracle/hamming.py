"""cnn_vtl distance oracle. Follows src/cnn_vtl/similarity/DistanceCalculator.py:4-12. PINNED (tests/golden).
bin(a ^ b).count('1') on numpy int8: the XOR is an int8, bin() of a negative is '-0b…' of its magnitude, so the
count is popcount(|int8(a ^ b)|)  (|-128| = 128 -> 1 bit)."""
import numpy as np

_POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.int64)
# LUT indexed by the XOR byte reinterpreted as unsigned
SIGNED_LUT = np.array([bin(int(np.int8(np.uint8(i).view(np.int8)))).count("1") for i in range(256)], dtype=np.int64)


def distance(d1, d2, signed_bin_quirk=True):
    x = (np.asarray(d1, dtype=np.int8) ^ np.asarray(d2, dtype=np.int8)).view(np.uint8)
    return int((SIGNED_LUT if signed_bin_quirk else _POP8)[x].sum())


def distance_matrix(desc, signed_bin_quirk=True):
    """Full N x N matrix incl. diagonal (src/cnn_vtl/create_distance_matrix.py:31-36)."""
    desc = np.asarray(desc, dtype=np.int8)
    lut = SIGNED_LUT if signed_bin_quirk else _POP8
    n = len(desc)
    out = np.empty((n, n), dtype=np.int64)
    for i in range(n):
        out[i] = lut[(desc ^ desc[i]).view(np.uint8)].sum(axis=1)
    return out

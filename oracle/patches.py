"""Patch extraction oracle. Follows src/sdav/input/CvInputParser.py (reference) line by line."""
import numpy as np


def round_half_even(v):
    """Python's round() on a float (CvInputParser.py:111): round half to even."""
    return np.rint(np.asarray(v, dtype=np.float64)).astype(np.int64)


def window_bounds(length, centers, patch_size):
    """get_1d_boundaries (CvInputParser.py:49-89): [lo, hi] of a patch_size window around each centre, shifted
    forward when it starts below 0 and back when it ends beyond length-1."""
    if patch_size % 2 == 0:
        raise ValueError("Invalid patch size. Patch size must be an odd number")  # :64-65
    half = patch_size // 2
    centers = np.asarray(centers, dtype=np.int64)
    lo = centers - half
    hi = centers + half
    shift_forward = (lo < 0) * lo * -1          # :80
    aux = hi - length + 1                        # :81
    shift_back = (aux > 0) * aux                 # :82
    return lo - shift_back + shift_forward, hi - shift_back + shift_forward


def extract_patches(img, xy, patch_size=41, swap_xy_quirk=True):
    """get_vectorized_patches_from_key_points + `/ 255.0` (CvInputParser.py:100-123, 27).

    img: uint8 [H, W]; xy: float [P, 2] keypoint (x, y) centres. Returns float64 [P, patch_size**2].
    With swap_xy_quirk (reference behaviour) rows are indexed by x (bounded by H) and columns by y (bounded by W)
    (CvInputParser.py:92-97, 119)."""
    img = np.asarray(img)
    H, W = img.shape
    xy = np.asarray(xy)
    x = round_half_even(xy[:, 0])
    y = round_half_even(xy[:, 1])
    if swap_xy_quirk:
        r_lo, r_hi = window_bounds(H, x, patch_size)
        c_lo, c_hi = window_bounds(W, y, patch_size)
    else:
        r_lo, r_hi = window_bounds(H, y, patch_size)
        c_lo, c_hi = window_bounds(W, x, patch_size)
    out = np.empty((len(xy), patch_size ** 2), dtype=np.int64)
    for i in range(len(xy)):
        out[i] = img[r_lo[i]:r_hi[i] + 1, c_lo[i]:c_hi[i] + 1].reshape(patch_size ** 2)
    return out / 255.0

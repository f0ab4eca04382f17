"""Global matcher oracle (cosine / dot / squared-L2 similarity matrix + per-row top-k / threshold).
NOT IN THE REFERENCE (the north star adds it) - this file is the definition; parity unpinned.
Order: best first, ties -> lowest index (stable argsort)."""
import numpy as np

COS, DOT, L2 = 0, 1, 2


def normalise_rows(x):
    x = np.asarray(x, dtype=np.float64)
    n = np.linalg.norm(x, axis=1, keepdims=True)
    return np.divide(x, n, out=np.zeros_like(x), where=n > 0)


def score_matrix(q, db_stored, metric):
    """q float [B, D]; db_stored: the database rows AS STORED (already normalised + rounded for COS)."""
    q = np.asarray(q, dtype=np.float64)
    d = np.asarray(db_stored, dtype=np.float64)
    if metric == COS:
        return normalise_rows(q) @ d.T
    if metric == DOT:
        return q @ d.T
    return (q * q).sum(1)[:, None] + (d * d).sum(1)[None, :] - 2.0 * (q @ d.T)


def topk(scores, k, largest=True, exclude_band=-1):
    s = np.array(scores, dtype=np.float64)
    rows, cols = s.shape
    out_s = np.full((rows, k), -np.inf if largest else np.inf)
    out_i = np.full((rows, k), -1, dtype=np.int64)
    for r in range(rows):
        cand = np.arange(cols)
        if exclude_band >= 0:
            cand = cand[np.abs(cand - r) > exclude_band]
        v = s[r, cand]
        keep = ~np.isnan(v)
        cand, v = cand[keep], v[keep]
        order = np.argsort(-v if largest else v, kind="stable")[:k]
        out_s[r, :len(order)] = v[order]
        out_i[r, :len(order)] = cand[order]
    return out_s, out_i


def threshold(scores, thr, max_per_row, smaller_is_better=False):
    s = np.asarray(scores, dtype=np.float64)
    passed = s <= thr if smaller_is_better else s >= thr
    counts = passed.sum(1).astype(np.int32)
    ts, ti = topk(s, max_per_row, largest=not smaller_is_better)
    ok = (ts <= thr) if smaller_is_better else (ts >= thr)
    ok &= ti >= 0
    ts = np.where(ok, ts, np.inf if smaller_is_better else -np.inf)
    ti = np.where(ok, ti, -1)
    return counts, ts, ti

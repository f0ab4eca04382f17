"""From-pixels parity report (test infrastructure): compares the scores of a device pipeline (pixels -> patches ->
encoder -> SDAV scores) with the float64 oracle run on the SAME pixels, and explains every score outside the tolerance.

The north star's bar: descriptors and scores within 1e-3 relative, candidate lists identical except for ties within
that tolerance, ties reported. A from-pixels comparison differs from a stage-wise one: the device descriptors carry
their own (bounded) error, the reference's matcher takes a hard arg-min over 30 patches (SimilarityCalculator.py:29-37)
and a logarithm of a difference that can be arbitrarily small (:47-49), so a descriptor error far inside the tolerance
can legitimately flip a nearest neighbour or move b*ln(s_k). Every frame pair whose score is outside the tolerance is
therefore put in exactly one class:

  ok            |S_dev - S_ref| <= tol * max(1, |S_ref|)
  tie           some rows match a different patch than the oracle's, and for each of them the device's choice is
                within the MEASURED descriptor error of the oracle's nearest neighbour:
                    ||a - g|| - ||a - o|| <= 2 ||da|| + ||dg|| + ||do||      (triangle inequality; a, g, o oracle rows,
                d* = device row - oracle row), and the device score equals the float64 score of the device's own
                descriptors (the matcher itself is exact)
  conditioning  same matches everywhere, but sum_k |b| (|dp_a| + |dp_b|) / min(s_k) - the first-order propagation of
                the measured descriptor error through b*ln|p_a - p_b| - covers the difference
  unexplained   anything else: a parity failure
"""
import warnings

import numpy as np

from oracle import similarity as o_sim


def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return np.abs(got - want) / np.maximum(1.0, np.abs(want))


def _dist(a, rows):
    return np.linalg.norm(rows - a, axis=1)


def classify_pair(s_dev, hd1, hd2, ho1, ho2, w_dev, w_ref, a=10.0, b=-10.0, tol=1e-3, s_ref=None, idx_ref=None):
    """Class of one ordered frame pair. hd*: device descriptors [P, D] (float64 copies), ho*: oracle descriptors;
    w_dev / w_ref: distinctive weights of the respective descriptor sets. Returns (class, details)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if s_ref is None or idx_ref is None:
            s_ref, idx_ref, sk_ref = o_sim.similarity_score(ho1, ho2, w_ref, a, b, return_details=True)
        else:
            sk_ref = np.abs((ho1 - ho2[idx_ref]) @ w_ref)
        if not np.isfinite(s_ref) and not np.isfinite(s_dev):
            return "ok", {}
        if abs(s_dev - s_ref) <= tol * max(1.0, abs(s_ref)):
            return "ok", {}
        stage, idx_dev, sk_dev = o_sim.similarity_score(hd1, hd2, w_dev, a, b, return_details=True)
    det = {"s_dev": float(s_dev), "s_ref": float(s_ref), "s_stage": float(stage)}
    if not (abs(s_dev - stage) <= tol * max(1.0, abs(stage))):
        det["why"] = "device score differs from the float64 score of the device's own descriptors"
        return "unexplained", det
    flips = np.nonzero(idx_dev != idx_ref)[0]
    if len(flips):
        worst = 0.0
        for k in flips:
            g, o = int(idx_dev[k]), int(idx_ref[k])
            lhs = np.linalg.norm(ho1[k] - ho2[g]) - np.linalg.norm(ho1[k] - ho2[o])
            rhs = (2 * np.linalg.norm(hd1[k] - ho1[k]) + np.linalg.norm(hd2[g] - ho2[g]) +
                   np.linalg.norm(hd2[o] - ho2[o]))
            worst = max(worst, lhs - rhs)
            if lhs > rhs * (1 + 1e-9) + 1e-12:
                det["why"] = "row %d matches patch %d instead of %d, outside the measured descriptor error" % (k, g, o)
                det["excess"] = float(lhs - rhs)
                return "unexplained", det
        det["flipped_rows"] = int(len(flips))
        return "tie", det
    # same matches: first-order propagation of the measured descriptor (and weight) error through b * ln|p_a - p_b|
    dp = np.abs(hd1 @ w_dev - ho1 @ w_ref) + np.abs((hd2 @ w_dev - ho2 @ w_ref)[idx_ref])
    smin = np.minimum(sk_dev, sk_ref)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(smin > 0, dp / smin, np.where(dp > 0, np.inf, 0.0))   # s_k = 0: identical matched patches
        bound = float(np.sum(abs(b) * ratio))
    det["bound"] = bound
    if abs(stage - s_ref) <= 2.0 * bound + tol * max(1.0, abs(s_ref)):
        return "conditioning", det
    det["why"] = "same matches, difference beyond the propagated descriptor error"
    return "unexplained", det


def report(S_dev, desc_dev, desc_ref, S_ref=None, idx_ref=None, full_asymmetric=True, a=10.0, b=-10.0, tol=1e-3):
    """S_dev [N, N] device scores; desc_dev / desc_ref [N, P, D]; S_ref / idx_ref: optional precomputed oracle scores
    [N, N] and matches [N, N, P] (fixtures). Returns a dict of counts plus the descriptor error."""
    dd = np.asarray(desc_dev, dtype=np.float64)
    dr = np.asarray(desc_ref, dtype=np.float64)
    n = dd.shape[0]
    w_dev = o_sim.distinctive_weights(dd)
    w_ref = o_sim.distinctive_weights(dr)
    counts = {"ok": 0, "tie": 0, "conditioning": 0, "unexplained": 0}
    worst_ok = 0.0
    bad = []
    for i in range(n):
        for j in range(n):
            if i == j or (not full_asymmetric and j < i):
                continue
            s_ref = None if S_ref is None else float(S_ref[i, j])
            ir = None if idx_ref is None else np.asarray(idx_ref[i, j], dtype=np.int64)
            cls, det = classify_pair(float(S_dev[i, j]), dd[i], dd[j], dr[i], dr[j], w_dev, w_ref, a, b, tol, s_ref, ir)
            counts[cls] += 1
            if cls == "ok" and s_ref is not None and np.isfinite(s_ref):
                worst_ok = max(worst_ok, float(rel_err(S_dev[i, j], s_ref)))
            if cls == "unexplained":
                bad.append(((i, j), det))
    out = dict(counts)
    out["pairs"] = int(sum(counts.values()))
    out["descriptor_max_rel_err"] = float(rel_err(dd, dr).max())
    out["descriptor_normwise_rel_err"] = float(np.linalg.norm(dd - dr) / np.linalg.norm(dr))
    out["worst_ok_score_rel_err"] = worst_ok
    out["unexplained_detail"] = bad[:10]
    return out


def pair_classes(S_dev, desc_dev, desc_ref, S_ref, idx_ref=None, a=10.0, b=-10.0, tol=1e-3):
    """[N, N] array of class names for the mirrored i<j matrix (create_similarity_matrix.py:31-38)."""
    dd = np.asarray(desc_dev, dtype=np.float64)
    dr = np.asarray(desc_ref, dtype=np.float64)
    n = dd.shape[0]
    w_dev = o_sim.distinctive_weights(dd)
    w_ref = o_sim.distinctive_weights(dr)
    cls = np.full((n, n), "ok", dtype=object)
    for i in range(n):
        for j in range(i + 1, n):
            ir = None if idx_ref is None else np.asarray(idx_ref[i, j], dtype=np.int64)
            c, _ = classify_pair(float(S_dev[i, j]), dd[i], dd[j], dr[i], dr[j], w_dev, w_ref, a, b, tol,
                                 float(S_ref[i, j]), ir)
            cls[i, j] = cls[j, i] = c
    return cls


def candidate_report(idx_dev, S_ref, cls, k, tol=1e-3):
    """Loop-candidate lists (per-row top-k of the mirrored score matrix, frame itself excluded, best first, ties ->
    lowest index) of the device against the oracle's ranking. A row's device list is checked through its INVERSIONS:
    pairs (d, o) with d listed by the device, o ranked strictly better than d by the oracle, and o NOT listed before
    d by the device (o is missing from the list or comes later). Comparing position by position would blame every
    entry behind one moved candidate. An inversion is a
      tie within tolerance  when both frame pairs are class ok and the oracle scores differ by at most 2 tol (each
                            device score may be off by tol),
      moved pair            when (row, d) or (row, o) is a tie / conditioning pair (its score legitimately moved),
      unexplained           otherwise."""
    S = np.array(S_ref, dtype=np.float64)
    n = len(S)
    np.fill_diagonal(S, -np.inf)
    want = np.argsort(-S, axis=1, kind="stable")[:, :k]
    out = {"rows": n, "rows_identical": 0, "rows_with_inversions": 0, "inversions_tie_within_tol": 0,
           "inversions_moved_pair": 0, "positions_unexplained": 0}
    for r in range(n):
        if np.array_equal(idx_dev[r], want[r]):
            out["rows_identical"] += 1
            continue
        out["rows_with_inversions"] += 1
        listed_before = np.zeros(n, dtype=bool)
        for p in range(k):
            d = int(idx_dev[r, p])
            if d < 0:
                continue
            better = np.nonzero((S[r] > S[r, d]) & ~listed_before)[0]      # the oracle prefers these to d
            for o in better:
                if cls[r, d] != "ok" or cls[r, o] != "ok":
                    out["inversions_moved_pair"] += 1
                elif S[r, o] - S[r, d] <= 2 * tol * max(1.0, abs(S[r, o])):
                    out["inversions_tie_within_tol"] += 1
                else:
                    out["positions_unexplained"] += 1
            listed_before[d] = True
    return out

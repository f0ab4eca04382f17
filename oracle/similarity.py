"""SDAV similarity oracle. Follows src/sdav/similarity/SimilarityCalculator.py:12-49 and the i<j loop of
src/sdav/create_similarity_matrix.py:31-38. PINNED against the imported reference (tests/golden)."""
import numpy as np


def distinctive_weights(dataset, mu=0.5, sigma=0.2):
    """_average_response + _distinctive_score (SimilarityCalculator.py:19-27)."""
    d = np.asarray(dataset, dtype=np.float64)
    avg = np.average(d.reshape(d.shape[0] * d.shape[1], d.shape[2]), axis=0)
    return np.exp(-((avg - mu) ** 2) / (2 * sigma ** 2))


def match_indices(h1, h2):
    """_match_features (:29-37): for every row of h1 the index of the first L2-nearest row of h2."""
    idx = np.empty(len(h1), dtype=np.int64)
    for k, mk in enumerate(h1):
        idx[k] = np.argmin(np.linalg.norm(h2 - mk, axis=1))
    return idx


def similarity_score(h1, h2, w, a=10.0, b=-10.0, return_details=False):
    """similarity_score with the dataset weights hoisted (:12-17, 39-49)."""
    h1 = np.asarray(h1, dtype=np.float64)
    h2 = np.asarray(h2, dtype=np.float64)
    idx = match_indices(h1, h2)
    s = np.abs((h1 - h2[idx]) @ w)                # |w . (h1[k] - h2[j*])|  (np.matmul then norm of a scalar)
    with np.errstate(divide="ignore"):
        score = np.sum(a + b * np.log(s))
    return (score, idx, s) if return_details else score


def similarity_matrix(dataset, mu=0.5, sigma=0.2, a=10.0, b=-10.0, full_asymmetric=False, return_details=False):
    """All pairs. Default = reference behaviour: i<j evaluated, mirrored, diagonal -1 (create_similarity_matrix.py
    :31-38) but kept in float (the reference's int64 matrix truncates on store)."""
    d = np.asarray(dataset, dtype=np.float64)
    n = d.shape[0]
    w = distinctive_weights(d, mu, sigma)
    S = np.full((n, n), -1.0)
    det = {}
    for i in range(n):
        for j in range(n):
            if i == j or (not full_asymmetric and j < i):
                continue
            sc, idx, s = similarity_score(d[i], d[j], w, a, b, return_details=True)
            S[i, j] = sc
            if not full_asymmetric:
                S[j, i] = sc
            if return_details:
                det[(i, j)] = (idx, s)
    return (S, det) if return_details else S


def nn_margin(h1, h2):
    """Gap between the best and second-best squared distance for every row of h1 (tie diagnostics)."""
    d2 = ((h1[:, None, :] - h2[None, :, :]) ** 2).sum(-1)
    part = np.partition(d2, 1, axis=1)
    return part[:, 1] - part[:, 0]

"""Matrix -> image oracle: the tails of src/sdav/create_similarity_matrix.py:31,41-48 and
src/cnn_vtl/create_distance_matrix.py:31,40-41. TEST INFRASTRUCTURE ONLY.
PINNED (tests/golden/images.npz): the fixtures are produced by the reference's own lines run literally on NumPy
followed by the real cv2.imwrite -> cv2.imread round trip (tests/golden/make_golden.py)."""
import numpy as np


def to_reference_int(scores):
    """np.full([n, n], -1) is an int64 matrix (:31): storing a float score truncates it toward zero."""
    return np.trunc(np.asarray(scores, dtype=np.float64)).astype(np.int64)


def similarity_image_f64(similarity_matrix):
    """create_similarity_matrix.py:41-45, operation order kept."""
    m = np.asarray(similarity_matrix)
    move_factor = 0 - m.min()
    divide_factor = m.max() + move_factor
    normalized_matrix = (m + move_factor) / divide_factor
    return 255 * normalized_matrix


def distance_image_f64(distance_matrix):
    """create_distance_matrix.py:40."""
    m = np.asarray(distance_matrix)
    return 255 - m / m.max() * 255


def imwrite_u8(img_f64):
    """cv2.imwrite on a float64 array converts with saturate_cast<uchar>(cvRound(x)): round half to even, clamp."""
    with np.errstate(invalid="ignore"):
        r = np.rint(np.asarray(img_f64, dtype=np.float64))
    r = np.where(np.isnan(r), 0.0, r)
    return np.clip(r, 0, 255).astype(np.uint8)

"""Drop-in for `src.cnn_vtl.similarity.DistanceCalculator.DistanceCalculator`
(reference src/cnn_vtl/similarity/DistanceCalculator.py:4-12) plus the N x N matrix of
src/cnn_vtl/create_distance_matrix.py:31-36. Integer-exact on the B200 (XOR + popcount kernel, including the
reference's signed-`bin()` semantics)."""
import numpy as np


class DistanceCalculator:
    @staticmethod
    def calculate_distance(desc1, desc2, signed_bin_quirk=True):
        d = DistanceCalculator.distance_matrix(np.stack([np.asarray(desc1), np.asarray(desc2)]), signed_bin_quirk)
        return int(d[0, 1])

    @staticmethod
    def distance_matrix(descriptors, signed_bin_quirk=True):
        import torch

        from . import _cuda, ops
        _cuda.require_cuda()
        d = np.asarray(descriptors)
        if d.dtype != np.int8:
            if np.any(d < -128) or np.any(d > 127):
                raise ValueError("descriptors must fit int8")
            d = d.astype(np.int8)
        return ops.hamming_matrix(torch.from_numpy(np.ascontiguousarray(d)).cuda(), signed_bin_quirk).cpu().numpy()


def distance_image(distance_matrix):
    """Hamming matrix -> uint8 image on the B200: 255 - D / max * 255 (create_distance_matrix.py:40) and cv2.imwrite's
    float64 -> uint8 conversion (:41). Accepts a NumPy array or a CUDA int32 tensor; returns NumPy."""
    import torch

    from . import _cuda, ops
    _cuda.require_cuda()
    m = distance_matrix
    if not isinstance(m, torch.Tensor):
        m = np.asarray(m)
        if m.size and (m.min() < -2 ** 31 or m.max() >= 2 ** 31):
            raise ValueError("distances must fit int32")
        m = torch.from_numpy(np.ascontiguousarray(m, dtype=np.int32))
    m = m.to(device="cuda", dtype=torch.int32).contiguous()
    return ops.matrix_image(m, ops.IMG_DISTANCE).cpu().numpy()

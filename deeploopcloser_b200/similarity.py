"""Drop-in for `src.sdav.similarity.SimilarityCalculator.SimilarityCalculator`
(reference src/sdav/similarity/SimilarityCalculator.py:4-49) plus the batched matrix the reference builds with a
Python double loop (src/sdav/create_similarity_matrix.py:31-38). All arithmetic runs on the B200 (Gram + argmin + score
tcgen05 kernel)."""
import numpy as np


class SimilarityCalculator:
    def __init__(self, dataset, mu=0.5, sigma=0.2, a=10, b=-10, precision="auto"):
        self.mu = mu
        self.sigma = sigma
        self.a = a
        self.b = b
        self.dataset = dataset
        self.precision = precision
        self._desc = None
        self._w = None

    def _device_dataset(self):
        import torch

        from . import _cuda, ops
        if self._desc is None:
            _cuda.require_cuda()
            d = np.asarray(self.dataset)
            if d.ndim != 3:
                raise ValueError("dataset must have shape [N, patches, features]")
            self._desc = torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32)).cuda()
            self._w = ops.sdav_weights(self._desc, self.mu, self.sigma)
        return self._desc, self._w

    def similarity_score(self, h1, h2):
        """Score of frame h1 against frame h2 ([P, D] each), with the distinctive weights of `self.dataset`."""
        import torch

        from . import ops
        _, w = self._device_dataset()
        pair = np.stack([np.asarray(h1, dtype=np.float32), np.asarray(h2, dtype=np.float32)])
        S = ops.sdav_similarity(torch.from_numpy(pair).cuda(), self.mu, self.sigma, self.a, self.b, weights=w,
                                precision=self.precision, full_asymmetric=True)
        return float(S[0, 1].item())

    def similarity_matrix(self, full_asymmetric=False, reference_int=False):
        """All frame pairs of the dataset -> [N, N]. Default order = the reference script: i<j evaluated, mirrored,
        diagonal -1. reference_int=True truncates toward zero like the reference's int64 matrix
        (create_similarity_matrix.py:31)."""
        from . import ops
        desc, w = self._device_dataset()
        S = ops.sdav_similarity(desc, self.mu, self.sigma, self.a, self.b, weights=w, precision=self.precision,
                                full_asymmetric=full_asymmetric).cpu().numpy()
        if reference_int:
            with np.errstate(invalid="ignore"):
                return np.trunc(S).astype(np.int64)
        return S

    def loop_candidates(self, k=10, exclude_band=0):
        """Per-frame top-k most similar frames (new; north star): (scores [N,k], indices [N,k])."""
        from . import ops
        desc, w = self._device_dataset()
        S = ops.sdav_similarity(desc, self.mu, self.sigma, self.a, self.b, weights=w, precision=self.precision)
        s, i = ops.topk_rows(S, k, largest=True, exclude_band=exclude_band)
        return s.cpu().numpy(), i.cpu().numpy()


def similarity_image(similarity_matrix, reference_int=True):
    """Score matrix -> uint8 image on the B200: min-max normalisation to 0..255 (create_similarity_matrix.py:41-45)
    and cv2.imwrite's float64 -> uint8 conversion (:48). reference_int=True first truncates the scores toward zero,
    as the reference's int64 matrix does on store (:31). Accepts a NumPy array or a CUDA tensor; returns NumPy."""
    import torch

    from . import _cuda, ops
    _cuda.require_cuda()
    m = similarity_matrix
    if not isinstance(m, torch.Tensor):
        m = torch.from_numpy(np.ascontiguousarray(m, dtype=np.float32))
    m = m.to(device="cuda", dtype=torch.float32).contiguous()
    return ops.matrix_image(m, ops.IMG_SIMILARITY, truncate_int=reference_int).cpu().numpy()


def write_png(path, image_u8):
    """8-bit grey PNG (what cv2.imwrite produces for the reference's matrices); pure zlib, no OpenCV needed."""
    import struct
    import zlib
    img = np.ascontiguousarray(image_u8, dtype=np.uint8)
    if img.ndim != 2:
        raise ValueError("expected a [rows, cols] uint8 image")
    h, w = img.shape

    def chunk(tag, data):
        body = tag + data
        return struct.pack(">I", len(data)) + body + struct.pack(">I", zlib.crc32(body) & 0xFFFFFFFF)

    raw = b"".join(b"\x00" + img[r].tobytes() for r in range(h))          # filter type 0 per scanline
    png = (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) +
           chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
    with open(str(path), "wb") as f:
        f.write(png)

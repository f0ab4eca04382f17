"""Drop-in for the reference script src/sdav/create_similarity_matrix.py: encode every frame of a dataset with the
SDA patch encoder, score all frame pairs (i < j, mirrored, diagonal -1, int64 truncation of :31), normalise to
0..255 and write the PNG (:41-48) - every stage on the B200. The reference hard-codes its paths; here they are
arguments:

    python -m src.sdav.create_similarity_matrix DATASET_DIR OUT.png [--checkpoint DIR] [--keypoints seeded|surf]

`--keypoints surf` (default) finds the 30 strongest fast-Hessian keypoints of every frame with the B200 detector
(dlc_surf_detect - the reference calls cv2.xfeatures2d.SURF_create().detect, CvInputParser.py:36-46; OpenCV's non-free
module is not needed and not used). `--keypoints seeded` is an explicit opt-in for tests: 30 uniform keypoints per
frame from np.random.default_rng(seed)."""
import argparse
import glob
import logging
import os

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("dataset")
    ap.add_argument("out_png")
    ap.add_argument("--checkpoint", default=None, help="SDAV checkpoint directory (TensorFlow format); default: N(0,1) init")
    ap.add_argument("--keypoints", default="surf", choices=["seeded", "surf"])
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    logging.getLogger().setLevel(logging.INFO)

    import cv2

    from src.sdav.network.SDAV import SDAV
    from src.sdav.similarity.SimilarityCalculator import SimilarityCalculator
    from deeploopcloser_b200.similarity import similarity_image, write_png

    network = SDAV()
    if args.checkpoint:
        network.load_weights(args.checkpoint)
    files = sorted(glob.glob(os.path.join(args.dataset, "*")))
    key_points = None                # "surf": CvInputParser runs the device detector on every frame
    if args.keypoints == "seeded":
        rng = np.random.default_rng(args.seed)

        def key_points(i, path):
            h, w = cv2.imread(path, cv2.IMREAD_GRAYSCALE).shape
            return np.stack([rng.uniform(0, w, network.input_shape[0]), rng.uniform(0, h, network.input_shape[0])], 1)

    logging.info("Transforming parsed dataset into descriptors")
    descriptors = network.transform_all(os.path.join(args.dataset, "*"), key_points=key_points)
    logging.info("Calculating similarity")
    calculator = SimilarityCalculator(np.array(descriptors))
    similarity_matrix = calculator.similarity_matrix(reference_int=True)
    write_png(args.out_png, similarity_image(similarity_matrix))
    logging.info("Wrote %s (%d x %d)", args.out_png, len(files), len(files))
    return similarity_matrix


if __name__ == "__main__":
    main()

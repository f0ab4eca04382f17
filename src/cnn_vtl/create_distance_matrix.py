"""Drop-in for the reference script src/cnn_vtl/create_distance_matrix.py: read every frame of a dataset (sorted),
run the cnn_vtl conv head, fill the full N x N Hamming matrix (:31-36), map it to 255 - D / max * 255 and write the
PNG (:40-41) - every stage on the B200. Paths are arguments instead of the reference's hard-coded ones:

    python -m src.cnn_vtl.create_distance_matrix DATASET_DIR OUT.png [--weights bvlc_alexnet.npy] [--seed 0]"""
import argparse
import logging
import os

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("dataset")
    ap.add_argument("out_png")
    ap.add_argument("--weights", default=None, help="bvlc_alexnet.npy; default: seeded He-scaled weights")
    ap.add_argument("--seed", type=int, default=0, help="seed of the column mask (and of the synthetic weights)")
    args = ap.parse_args(argv)
    logging.getLogger().setLevel(logging.INFO)

    import cv2

    from src.cnn_vtl.network.cnn_vtl import CnnVtl
    from src.cnn_vtl.similarity.DistanceCalculator import DistanceCalculator
    from deeploopcloser_b200.distance import distance_image
    from deeploopcloser_b200.similarity import write_png

    files = sorted(os.listdir(args.dataset))
    n_files = len(files)
    logging.info("Reading files...")
    dataset = [cv2.imread(os.path.join(args.dataset, f)) for f in files]
    h, w = dataset[0].shape[:2]
    logging.info("Creating network with shape=[%d, %d, %d, 3]" % (n_files, h, w))
    network = CnnVtl(input_shape=[n_files, h, w, 3], weights=args.weights or "synthetic", seed=args.seed)
    logging.info("Transforming images into descriptors...")
    descriptors = network.transform(np.stack(dataset))
    logging.info("Creating distance matrix...")
    distance_matrix = DistanceCalculator.distance_matrix(descriptors)
    write_png(args.out_png, distance_image(distance_matrix))
    logging.info("Wrote %s (%d x %d)", args.out_png, n_files, n_files)
    return distance_matrix


if __name__ == "__main__":
    main()

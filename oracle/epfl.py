"""Oracle loaders for the EPFL triplet data (TEST INFRASTRUCTURE).

Follows Data/readCalibrationOrientation_EPFL.m:5-22 and the per-triplet
preparation of experiments_real.m:78-109.  The data files themselves stay in
/root/reference (never copied); tests that need them skip when it is absent.
"""
import os

import numpy as np

from .reference_port import triangulation3D, project3Dpoints, ReprError


def _nums(line):
    return [float(t) for t in line.split()]


def readCalibrationOrientation_EPFL(image_path, image_name):
    """Data/readCalibrationOrientation_EPFL.m:5-22 -> K, R, t, im_size."""
    filename = os.path.join(image_path, image_name + '.camera')     # :5
    with open(filename, 'r') as f:
        lines = f.read().splitlines()
    K = np.array([_nums(lines[0]), _nums(lines[1]), _nums(lines[2])])  # :8-10
    # lines[3] skipped                                              # :12
    R = np.array([_nums(lines[4]), _nums(lines[5]), _nums(lines[6])]).T  # :14-16
    t = -R @ np.array(_nums(lines[7]))                              # :18
    im_size = np.array(_nums(lines[8]))                             # :20
    return K, R, t, im_size


def load_corresp_triplets(path_to_data):
    """experiments_real.m:45-48 (MAT v5 via scipy)."""
    import scipy.io as sio
    m = sio.loadmat(os.path.join(path_to_data, 'Corresp_triplets.mat'))
    im_names = [str(x[0]) for x in m['im_names'].ravel()]
    return m['indexes_sorted'].astype(np.int64), m['Corresp'], im_names


def epfl_triplet(path_to_data, indexes_sorted, corresp_by_triplet, im_names, it,
                 repr_err_th=1.0):
    """experiments_real.m:78-101 for 1-based triplet number `it`.
    Returns dict(CalM, R_t0, Corresp, Corresp_inliers, REr)."""
    im1, im2, im3 = (int(v) for v in indexes_sorted[it - 1, 0:3])   # :78-79
    Corresp = np.asarray(corresp_by_triplet[im1 - 1, im2 - 1, im3 - 1], dtype=np.float64).T  # :80
    K1, R1, t1, _ = readCalibrationOrientation_EPFL(path_to_data, im_names[im1 - 1])  # :86
    K2, R2, t2, _ = readCalibrationOrientation_EPFL(path_to_data, im_names[im2 - 1])  # :87
    K3, R3, t3, _ = readCalibrationOrientation_EPFL(path_to_data, im_names[im3 - 1])  # :88
    CalM = np.vstack([K1, K2, K3])                                  # :89
    R_t0 = [np.column_stack([R2 @ R1.T, t2 - R2 @ R1.T @ t1]),
            np.column_stack([R3 @ R1.T, t3 - R3 @ R1.T @ t1])]      # :90-91
    Ps = [K1 @ np.eye(3, 4), K2 @ R_t0[0], K3 @ R_t0[1]]
    Reconst0 = triangulation3D(Ps, Corresp)                         # :94
    Reconst0 = Reconst0[0:3, :] / Reconst0[3:4, :]                  # :95
    Corresp_new = project3Dpoints(Reconst0, Ps)                     # :96
    residuals = Corresp_new - Corresp                               # :97
    mask = np.sum(np.abs(residuals) > repr_err_th, axis=0) == 0     # :98
    Corresp_inliers = Corresp[:, mask]
    REr = ReprError(Ps, Corresp_inliers)                            # :100
    return dict(CalM=CalM, R_t0=R_t0, Corresp=Corresp, Corresp_inliers=Corresp_inliers,
                inlier_mask=mask, REr=REr, triplet=(im1, im2, im3))

"""CPU oracle for the linear three-view pose path of LauraFJulia/TFT_vs_Fund.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product path (``tft_vs_fund_b200``) never imports this package and fails
loudly when its CUDA library is missing.

What it is: a line-by-line NumPy float64 restatement of the reference's MATLAB
functions on the hot path.  Every function cites the
reference file:line it follows.  The dense numerical built-ins the reference
gets from the MATLAB runtime (``svd``, ``inv``, ``rank``, ``norm``, ``det``,
``kron``) are taken from NumPy/LAPACK (``numpy.linalg.svd`` with
``full_matrices=True`` mirrors MATLAB's full ``svd``), so the oracle's
numerical route (bidiagonal SVD of the design matrix) is independent of the
GPU's (Gram matrix + inverse iteration / one-sided Jacobi).

Pin status: the reference ships no tests, golden vectors or stored outputs, and neither MATLAB nor
GNU Octave exists in this image, so the reference cannot be run under its native runtime ->
**parity is unpinned against MATLAB's own numerics** (its LAPACK build, `randn`, `randsample`).
What pins the restatement instead (tests/test_oracle_pinned.py, tests/test_oracle_kat.py):
  1. ``oracle/mini_matlab.py`` parses and executes the reference's *unmodified* ``.m`` sources
     from /root/reference (a MATLAB-subset interpreter whose built-ins are NumPy/LAPACK); its
     outputs on 260 sweep trials, the example.m scene and 12 EPFL triplets are committed as the
     ``ref_*`` arrays of ``tests/golden/*.npz`` and the restatement reproduces them to rounding.
  2. derived known-answer tests: noise-free scenes give the closed-form
     ``TFT_from_P`` tensor, ground-truth poses and zero reprojection error; the EPFL fountain
     triplet reproduces the inlier count and ground-truth reprojection RMS that
     experiments_real.m:101 prints (1360 inliers, 0.2586 px).
"""
from .reference_port import (  # noqa: F401
    Normalize2Ddata, linearTFT, transform_TFT, R_t_from_TFT, recover_R_t_TFT,
    recover_R_t_F, LinearTFTPoseEstimation, triangulation3D, ReprError,
    linearF, LinearFPoseEstimation, TFT_from_P, crossM, AngError,
    project3Dpoints, matlab_svd, matlab_rank, LinearFError,
)
from .gauss_helmert_port import Gauss_Helmert, optimF, OptimFPoseEstimation, matlab_pinv, constraintsGH_F  # noqa: F401
from .scene import generateSyntheticScene, SceneRNG, experiments_subsample  # noqa: F401
from .epfl import readCalibrationOrientation_EPFL, load_corresp_triplets, epfl_triplet  # noqa: F401

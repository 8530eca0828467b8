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

Pin status: the reference ships no tests, golden vectors or stored outputs,
and neither MATLAB nor GNU Octave exists in this image, so the reference
cannot be executed directly -> **parity unpinned against MATLAB itself**.
What pins the restatement instead (see tests/ and DESIGN.md):
  1. ``oracle/mini_matlab.py`` executes the *unmodified* reference ``.m``
     sources from /root/reference with a minimal MATLAB-subset interpreter
     (NumPy built-ins); its outputs are committed as ``tests/golden/*.npz``
     and the restatement must reproduce them.
  2. derived known-answer tests: noise-free scenes give
     the closed-form ``TFT_from_P`` tensor, ground-truth poses and zero
     reprojection error; EPFL triplets reproduce the inlier counts and
     ground-truth reprojection RMS the reference script prints.
"""
from .reference_port import (  # noqa: F401
    Normalize2Ddata, linearTFT, transform_TFT, R_t_from_TFT, recover_R_t_TFT,
    recover_R_t_F, LinearTFTPoseEstimation, triangulation3D, ReprError,
    linearF, LinearFPoseEstimation, TFT_from_P, crossM, AngError,
    project3Dpoints, matlab_svd, matlab_rank, LinearFError,
)
from .scene import generateSyntheticScene, SceneRNG, experiments_subsample  # noqa: F401
from .epfl import readCalibrationOrientation_EPFL, load_corresp_triplets, epfl_triplet  # noqa: F401

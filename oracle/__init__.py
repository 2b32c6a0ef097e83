"""CPU oracle for the PCReg alignment hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain numpy/C FP64 restatement of the reference's MATLAB
arithmetic (files under /root/reference, cited per function).  It exists to
CHECK the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``pcreg_b200`` never does and fails loudly without its CUDA
library.

Parity status (see DESIGN.md section "Oracle"):
  * estimateTransform / quickTF / invertTF: PINNED by the reference's one
    self-contained known-answer vector (testTransformEstimation.m:2-14) and the
    analytic identities of testRANSAC.m:27-29,40-42 (tests/golden/kat_*.json).
  * AlignPoints* family, ransac, getLocalPoints, getSpacialHistogramDescriptors (+ histcn): restated line by line from the
    .m files; MATLAB built-ins (pca/eig/sort/rank/round) restated from their
    documentation.  MATLAB/Octave are not installed here, so these are
    "parity unpinned" beyond algebraic identities.
  * getMatches: the weighting lines restated from getMatches.m; matchFeatures (closed toolbox, 'Approximate' in the
    reference's drivers) restated from its documentation as the exhaustive search.  PARITY UNPINNED.
  * Nearest-neighbour step and the composed ICP: the reference has no such
    code; semantics = MATLAB knnsearch documentation
    (Euclidean, FP64, K=1, ties -> smallest index).  PARITY UNPINNED.
"""
from .primitives import (  # noqa: F401
    eul2rotm, quickTF, invertTF, pcRigidBodyTF, getLocalPoints, getLocalPoints_v2,
    matlab_round, matlab_rank, estimateTransform, calcDists, ransac, ransac_triplets, check_alignment,
)
from .align import (  # noqa: F401
    pca_eig, AlignPoints, AlignPoints_KNN, AlignPoints_knn, AlignPoints_weighted,
    AlignPoints_c, AlignPoints_KNN_c,
)
from .descriptors import getSpacialHistogramDescriptors, spatial_histogram_of, histogram_edges, histcn3, histcounts_bin  # noqa: F401
from .matching import getMatches, match_features_exhaustive, weight_descriptors  # noqa: F401
from .nn import nn_brute, nn_kdtree  # noqa: F401
from .icp import icp_single, icp_batch, ICP_PLAIN, ICP_KNN, ICP_WEIGHTED  # noqa: F401

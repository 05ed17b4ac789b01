"""scmerge_b200: B200-native RUV-III fit-and-adjust path of scMerge behind the reference's interface.

The arithmetic lives in `libscmerge_b200.so` (hand-written sm_100a CUDA, C ABI in include/scmerge_b200.h);
this package is the host-side mirror of the R closures.  There is no CPU fallback.
"""
from ._lib import LIB_PATH, NonFiniteError, NotConvergedWarning, ScmError, SingularError, lib  # noqa: F401
from .device import Context, DeviceArray, DeviceMatrix, MultiContext, bind_host_to_gpu, default_context, pinned_empty  # noqa: F401
from .ruv import (ExactParam, FullAlpha, IrlbaParam, RandomParam, RUVFit, aggregate_matrix, createChunk,  # noqa: F401
                  create_pseudoBulk, fastRUVIII, genes_by_name, getAdjustedMat, multi_fastRUVIII, pseudoRUVIII, replicate_keys, replicate_matrix,
                  scRUVg, scRUVIII, tological, eta_design_rows)

from .scmerge2 import DeviceCSC, cosineNorm, csc_to_device, pseudobulk_keys, scMerge2_core  # noqa: F401,E402

from .replicate import (centroidDist, cluster_centroid_dist, compute_dist_mat, compute_dist_res,  # noqa: F401,E402
                        computePCA_byHVGMarker, identifyCluster, kmeans, run_pca_scores)

__version__ = "0.3.0"

"""warpdemux_b200 — B200-native classification hot path of WarpDemuX.

Public surface mirrors the reference for this path only:
    warpdemux_b200.models.dtw_svm.DTW_SVM.predict      (warpdemux/models/dtw_svm.py:54-98)
    warpdemux_b200.parallel_distances.distance_matrix_to (warpdemux/parallel_distances.py:48-84)
    warpdemux_b200.sig_proc.*                           (warpdemux/sig_proc.py:394-605)
All compute is in libwdx_b200.so (hand-written CUDA for sm_100a, C ABI in
include/wdx_b200.h).  No CPU fallback.
"""
__version__ = "0.1.0"

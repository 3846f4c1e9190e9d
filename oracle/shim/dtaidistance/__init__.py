"""TEST-ONLY stand-in for dtaidistance 2.3.13 (absent from this image and from
/root/reference).  It exists so that the reference's own Python modules
(`warpdemux/parallel_distances.py`, `warpdemux/models/dtw_svm.py`,
`warpdemux/sig_proc.py`) import UNMODIFIED from /root/reference when
oracle/make_golden.py generates fixtures.  Only `dtw.distance_matrix` computes
anything; it is backed by oracle/wdx_oracle.c.  Never on the product path.
"""
__version__ = "2.3.13+wdx-oracle-shim"

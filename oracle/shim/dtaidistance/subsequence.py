class SubsequenceAlignment:  # imported by sig_proc.py:17 (tRNA path only)
    def __init__(self, *a, **k):
        raise NotImplementedError("dtaidistance shim: SubsequenceAlignment not restated")

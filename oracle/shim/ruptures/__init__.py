"""TEST-ONLY import stub (sig_proc.py:18 imports KernelCPD for the tRNA path)."""


class KernelCPD:
    def __init__(self, *a, **k):
        raise NotImplementedError("ruptures stub")

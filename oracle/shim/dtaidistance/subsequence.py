"""`dtaidistance.subsequence.SubsequenceAlignment` as far as the reference uses it
(sig_proc.py:288-312): constructed with (query, series, use_c=True), `.paths` assigned from outside,
then `_compute_matching()` and `best_match().segment`.  Restated from the published 2.3.13 source
(dtaidistance/subsequence/dtw.py); TEST-ONLY, PARITY UNPINNED (see oracle/wdx_oracle.c)."""
import numpy as np

from . import dtw


class SAMatch:
    def __init__(self, idx, alignment):
        self.idx = idx
        self.alignment = alignment

    @property
    def value(self):
        return self.alignment.matching[self.idx]

    @property
    def distance(self):
        return self.value * len(self.alignment.query)

    @property
    def segment(self):
        start = self.alignment.matching_function_startpoint(self.idx)
        end = self.alignment.matching_function_endpoint(self.idx)
        return [start, end]


class SubsequenceAlignment:
    def __init__(self, query, series, penalty=0.1, use_c=False):
        self.query = query
        self.series = series
        self.penalty = penalty
        self.paths = None
        self.matching = None
        self.use_c = use_c

    def _compute_matching(self):
        matching = self.paths[-1, :]
        if len(matching) > len(self.series):
            matching = matching[-len(self.series):]
        self.matching = np.array(matching) / len(self.query)

    def get_match(self, idx):
        return SAMatch(idx, self)

    def best_match(self):
        best_idx = np.argmin(self.matching)
        return self.get_match(best_idx)

    def matching_function_endpoint(self, idx):
        return idx

    def matching_function_startpoint(self, idx):
        real_idx = idx + 1
        path = dtw.best_path(self.paths, col=real_idx)
        return path[0][1]

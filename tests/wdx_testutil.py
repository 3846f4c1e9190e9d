"""Shared helpers for the test-suite (uniquely named so it cannot collide with
other `tests` packages on sys.path)."""
import numpy as np


def synth_fingerprints(sv, n, seed=0, sigma=0.35):
    """S1 of SURVEY.md §8(d): support vectors + gaussian noise."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, sv.shape[0], size=n)
    return sv[idx] + sigma * rng.standard_normal((n, sv.shape[1]))

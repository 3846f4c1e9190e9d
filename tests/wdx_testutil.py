"""Shared helpers for the test-suite (uniquely named so it cannot collide with
other `tests` packages on sys.path)."""
import numpy as np


def synth_fingerprints(sv, n, seed=0, sigma=0.35):
    """S1 of SURVEY.md §8(d): support vectors + gaussian noise."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, sv.shape[0], size=n)
    return sv[idx] + sigma * rng.standard_normal((n, sv.shape[1]))


def synth_adapter_signals(n, seed=2, width=9000, short_frac=0.0):
    """S4 of SURVEY.md §8(d): piecewise-constant adapter-like signals (pA,
    float32), NaN-padded rows like the reference's minibatches
    (file_proc.py:333-354), plus adapter boundaries.  Same generator as
    oracle/make_golden.py (the first 32 rows of seed 2 are the golden inputs).
    `short_frac` of the rows get a short adapter (few levels, short dwells) so
    that the window / distance parameters fall below their caps."""
    rng = np.random.default_rng(seed)
    sig = np.full((n, width), np.nan, dtype=np.float32)
    a0 = np.zeros(n, dtype=np.int64)
    a1 = np.zeros(n, dtype=np.int64)
    for r in range(n):
        short = rng.random() < short_frac if short_frac > 0 else False
        n_levels = 121 + int(rng.integers(0, 60))
        levels = rng.standard_normal(n_levels) * 12.0 + 80.0
        dwell = (3 if short else 9) + rng.geometric(1.0 / (5.0 if short else 25.0), size=n_levels)
        x = np.repeat(levels, dwell)
        x = x + rng.normal(0.0, 2.0, size=x.size)
        spikes = rng.random(x.size) < 0.01
        x[spikes] += rng.choice([-60.0, 60.0], size=int(spikes.sum()))
        lead = int(rng.integers(0, 300))
        tail = int(rng.integers(200, 1500))
        full = np.concatenate([rng.normal(110.0, 3.0, lead), x, rng.normal(95.0, 6.0, tail)])[:width]
        sig[r, : full.size] = full.astype(np.float32)
        a0[r] = lead
        a1[r] = min(lead + x.size, full.size)
    a1[0] = a0[0] + 40  # tiny adapter: segmentation fails
    if n > 1:
        a0[1] = 0       # adapter at the very start of the read
    return sig, a0, a1


def oracle_fingerprints(sig, a0, a1, **cfg):
    """Row-by-row CPU oracle over a NaN-padded minibatch -> (status, fpt, dwell, stats)."""
    from oracle import wdx_oracle as o

    n = sig.shape[0]
    nb = cfg.get("barcode_num_events", 25)
    status = np.zeros(n, dtype=np.int32)
    fpt = np.full((n, nb), np.nan)
    dwell = np.zeros((n, nb), dtype=np.int64)
    stats = np.full((n, 6), np.nan)
    keys = ["adapter_dt_med", "adapter_dt_mad", "adapter_event_mean", "adapter_event_std", "adapter_event_med",
            "adapter_event_mad"]
    for r in range(n):
        valid = sig[r][~np.isnan(sig[r])]
        st, f, d, s = o.fingerprint(valid, int(a0[r]), int(a1[r]), **cfg)
        status[r] = st
        if st == 0:
            fpt[r], dwell[r] = f, d
            stats[r] = [s[k] for k in keys]
    return status, fpt, dwell, stats

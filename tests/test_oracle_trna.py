"""Consensus-guided (tRNA) fingerprint, SURVEY.md §8f rank 3: the CPU oracle against the fixture the
reference's own `detect_results_to_fpt` produced (oracle/make_golden_trna.py; real adapter signals of
test_data/demux/4000_rna004.pod5 and synthetic consensus-shaped signals).  The sub-sequence
alignment (dtaidistance, absent) is the oracle's restatement on both sides — PARITY UNPINNED for
that piece; everything around it is checked against the reference's own code."""
import hashlib
import json
import os

import numpy as np
import pytest

from wdx_testutil import oracle_fingerprints_consensus, real_fixture_rows, synth_trna_signals

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden_trna():
    return np.load(os.path.join(ROOT, "tests", "golden", "fingerprint_trna.npz"))


def trna_cfg(g):
    return json.loads(str(g["cfg"]))


def check(g, prefix, status, fpt, dwell, stats, cons):
    assert np.array_equal(status, g[prefix + "status"])
    ok = status == 0
    assert np.array_equal(fpt[ok], g[prefix + "fpt"][ok]), "fingerprints must be bit-identical"
    assert np.array_equal(dwell[ok], g[prefix + "dwell"][ok])
    rep = ok | (status == 5)                      # the reference reports statistics for these
    assert np.array_equal(stats[rep], g[prefix + "stats"][rep])
    assert np.array_equal(cons[rep], g[prefix + "cons"][rep])


def test_oracle_matches_reference_real_reads(golden_trna, golden_real):
    g = golden_trna
    rows = real_fixture_rows(golden_real)
    out = oracle_fingerprints_consensus(rows, golden_real["adapter_start"], golden_real["adapter_end"], g["consensus"],
                                        detect_ok=golden_real["detect_ok"], **trna_cfg(g))
    check(g, "real_", *out)
    ok = out[0] == 0
    assert ok.sum() >= 400                        # the consensus really is found in real adapter signals
    assert 80 <= np.median(out[4][ok, 1]) <= 90


def test_oracle_matches_reference_synthetic(golden_trna):
    g = golden_trna
    sig, a0, a1 = synth_trna_signals(g["consensus"], 48, seed=5)
    assert hashlib.sha256(np.ascontiguousarray(sig).tobytes()).hexdigest() == str(g["syn_signals_sha256"])
    out = oracle_fingerprints_consensus(sig, a0, a1, g["consensus"], **trna_cfg(g))
    check(g, "syn_", *out)
    assert set(np.unique(out[0])) >= {0, 1, 3, 5}


def test_warping_paths_properties():
    """Identities of the restated psi-relaxed warping paths: a query embedded in a series is found
    at its position with distance 0; without relaxation the corner equals the plain DTW distance."""
    from oracle import wdx_oracle as o

    rng = np.random.default_rng(3)
    q = rng.standard_normal(84)
    for lead in (0, 7, 30):
        s = np.concatenate([rng.standard_normal(lead) + 5.0, q, rng.standard_normal(121 - 84 - lead) + 5.0])
        start, end = o.subsequence_best_match(q, s, 1.5, (5, 0, 40, 0))
        assert (start, end) == (lead, lead + 83)
    a, b = rng.standard_normal(25), rng.standard_normal(25)
    d, paths = o.warping_paths(a, b, 0.1, (0, 0, 0, 0))
    assert d == paths[-1, -1] == o.dtw_matrix(a[None], b[None], 0, 0.1)[0, 0]
    assert np.isinf(paths[0, 1:]).all() and np.isinf(paths[1:, 0]).all() and paths[0, 0] == 0

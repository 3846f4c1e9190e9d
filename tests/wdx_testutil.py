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


def oracle_fingerprints(sig, a0, a1, padded_rows=True, **cfg):
    """Row-by-row CPU oracle over a NaN-padded minibatch -> (status, fpt, dwell, stats).
    padded_rows=True hands the whole padded row to the oracle, as the reference's worker does
    (file_proc.py:418-428); False strips the padding first (the "signal without NaNs" contract of
    barcode_fpt_wrapper, file_proc.py:194 — what an explicit `sig_len` means on the GPU path)."""
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
        row = sig[r] if padded_rows else sig[r][~np.isnan(sig[r])]
        with np.errstate(all="ignore"):
            st, f, d, s = o.fingerprint(row, int(a0[r]), int(a1[r]), **cfg)
        status[r] = st
        if st == 0:
            fpt[r], dwell[r] = f, d
            stats[r] = [s[k] for k in keys]
    return status, fpt, dwell, stats


def real_fixture_rows(g):
    """Rebuild the float32 pA minibatch rows of tests/golden/real_rna004_WDX4.npz
    (oracle/make_golden_real.py): NaN everywhere except the stored adapter slice;
    pA = (adc + calibration_offset) * calibration_scale in float32."""
    n, m = g["adapter_start"].size, int(g["preload_size"])
    rows = np.full((n, m), np.nan, dtype=np.float32)
    off = g["adc_offsets"]
    for r in range(n):
        adc = g["adc"][off[r]:off[r + 1]]
        s = int(g["slice_start"][r])
        rows[r, s:s + adc.size] = (adc.astype(np.float32) + g["calibration_offset"][r]) * g["calibration_scale"][r]
    return rows


def cnn_golden_signals(gold):
    """Rebuild the float32 NaN-padded minibatch rows of tests/golden/cnn_detect_rna004.npz
    (pA = (adc + offset) * scale in float32, file_proc.py:241-262)."""
    m = int(gold["preload_size"])
    offs = gold["adc_offsets"]
    n = len(offs) - 1
    sig = np.full((n, m), np.nan, dtype=np.float32)
    for i in range(n):
        adc = gold["adc"][offs[i]:offs[i + 1]].astype(np.float32)
        sig[i, : adc.size] = (adc + gold["calibration_offset"][i]) * gold["calibration_scale"][i]
    return sig


def synth_trna_signals(consensus, n, seed=5, width=9000):
    """Consensus-shaped adapter signals for the consensus-guided (tRNA) fingerprint
    (BASELINE.json configs[3]): `lead` random events, the constant adapter region (the
    consensus levels, some events dropped or split, level noise), then the barcode events; pA scale,
    float32, NaN-padded rows, adapter boundaries.  Rows 0..4 are edge cases: tiny adapter, adapter at
    the start of the read, a long lead (consensus found too late -> "consensus query outlier"), too
    few barcode events (second segmentation fails), NaN padding inside the slice."""
    rng = np.random.default_rng(seed)
    consensus = np.asarray(consensus, dtype=np.float64)
    sig = np.full((n, width), np.nan, dtype=np.float32)
    a0 = np.zeros(n, dtype=np.int64)
    a1 = np.zeros(n, dtype=np.int64)
    for r in range(n):
        n_lead = int(rng.integers(0, 22))
        n_bc = int(rng.integers(27, 42))
        if r == 2:
            n_lead = 45
        if r == 3:
            n_bc = 12
        keep = rng.random(consensus.size) > 0.04
        z = consensus[keep] + 0.08 * rng.standard_normal(int(keep.sum()))
        split = rng.random(z.size) < 0.04
        z = np.repeat(z, 1 + split.astype(int))
        levels = np.concatenate([rng.standard_normal(n_lead), z, rng.standard_normal(n_bc) * 1.1]) * 12.0 + 80.0
        dwell = 10 + rng.geometric(1.0 / 20.0, size=levels.size)
        x = np.repeat(levels, dwell)
        x = x + rng.normal(0.0, 1.5, size=x.size)
        spikes = rng.random(x.size) < 0.005
        x[spikes] += rng.choice([-60.0, 60.0], size=int(spikes.sum()))
        lead = int(rng.integers(0, 300))
        tail = 40 if r == 4 else int(rng.integers(200, 1500))
        full = np.concatenate([rng.normal(110.0, 3.0, lead), x, rng.normal(95.0, 6.0, tail)])[:width]
        sig[r, : full.size] = full.astype(np.float32)
        a0[r] = lead
        a1[r] = min(lead + x.size, full.size)
    a1[0] = a0[0] + 40
    if n > 1:
        a0[1] = 0
    return sig, a0, a1


def oracle_fingerprints_consensus(sig, a0, a1, consensus, detect_ok=None, **cfg):
    """Row-by-row CPU oracle of the consensus-guided fingerprint over a NaN-padded minibatch ->
    (status, fpt, dwell, stats, cons[n,3])."""
    from oracle import wdx_oracle as o

    cfg = dict(cfg)
    cfg.pop("consensus_model", None)
    n = sig.shape[0]
    nb = int(cfg.get("barcode_num_events", (25, 25))[1])
    status = np.zeros(n, dtype=np.int32)
    fpt = np.full((n, nb), np.nan)
    dwell = np.zeros((n, nb), dtype=np.int64)
    stats = np.full((n, 6), np.nan)
    cons = np.zeros((n, 3), dtype=np.int32)
    keys = ["adapter_dt_med", "adapter_dt_mad", "adapter_event_mean", "adapter_event_std", "adapter_event_med",
            "adapter_event_mad"]
    for r in range(n):
        if detect_ok is not None and not detect_ok[r]:
            status[r] = o.FP_FAIL_DETECT
            continue
        with np.errstate(all="ignore"):
            st, f, d, s, c = o.fingerprint_consensus(sig[r], int(a0[r]), int(a1[r]), consensus, **cfg)
        status[r] = st
        if st == 0:
            fpt[r], dwell[r] = f, d
        if s:
            stats[r] = [s[k] for k in keys]
            cons[r] = c
    return status, fpt, dwell, stats, cons


def validate_golden_inputs(gold_val, gold_cnn):
    """Inputs of tests/golden/validate_rna004.npz (oracle/make_golden_validate.py): the real reads of the CNN
    fixture with the reference CNN's own predictions, then synthetic rows rebuilt from their seeds.
    Returns (list of float32 rows, full_lens int64[n], preds int64[n, 1 + k])."""
    from oracle import wdx_oracle_validate as ov

    n_real, n_syn, k = int(gold_val["n_real"]), int(gold_val["n_syn"]), int(gold_val["k"])
    real = cnn_golden_signals(gold_cnn)
    assert real.shape[0] == n_real
    rows = [real[i] for i in range(n_real)]
    preds = [gold_cnn["preds"][i] for i in range(n_real)]
    for s in range(n_syn):
        row, _fl, pr = ov.synthetic_case(s, stride=real.shape[1], k=k)
        rows.append(row)
        preds.append(pr)
    return rows, gold_val["full_lens"].astype(np.int64), np.array(preds, dtype=np.int64)


def pack_rows(rows, fill=np.nan):
    """Ragged float32 rows -> one NaN-padded [n, max_len] minibatch."""
    m = max(r.size for r in rows)
    out = np.full((len(rows), m), fill, dtype=np.float32)
    for i, r in enumerate(rows):
        out[i, : r.size] = r
    return out


def unpack_adc_planes(lo, hi):
    """Inverse of oracle/make_golden_real4000.py::planes: byte planes of the zig-zag coded int16 deltas -> int16 samples."""
    z = lo.astype(np.uint32) | (hi.astype(np.uint32) << 8)
    d = ((z >> 1).astype(np.int32) ^ -(z & 1).astype(np.int32)).astype(np.int16)
    return np.cumsum(d.astype(np.int64)).astype(np.int16)      # int16 wrap-around like the encoder's deltas


def real4000_rows(g, full=None):
    """Minibatch rows (float32 pA, NaN padded to preload_size) of tests/golden/real4000_rna004_WDX4.npz.
    full = the dict of tests/golden/_local/real4000_adc_rows.npz -> all 4000 reads, else the committed subset.
    Returns (read indices int64 [m], rows float32 [m, preload], adc int16 [m, preload] (filler -7), num_samples int64 [m])."""
    m = int(g["preload_size"])
    src = full if full is not None else g
    idx = np.arange(len(g["full_lengths"])) if full is not None else g["subset"]
    adc_cat = unpack_adc_planes(src["adc_lo"], src["adc_hi"])
    offs = src["adc_offsets"]
    rows = np.full((idx.size, m), np.nan, dtype=np.float32)
    adc = np.full((idx.size, m), -7, dtype=np.int16)
    num = np.zeros(idx.size, dtype=np.int64)
    for j, i in enumerate(idx):
        a = adc_cat[offs[j]:offs[j + 1]]
        adc[j, : a.size] = a
        num[j] = a.size
        rows[j, : a.size] = (a.astype(np.float32) + g["calibration_offset"][i]) * g["calibration_scale"][i]
    return idx.astype(np.int64), rows, adc, num

"""Result formats either side of the path: warpdemux_b200/io/results.py against what the
reference's own file_proc functions wrote / returned on the same inputs (oracle/make_golden_io.py ->
tests/golden/io_formats.npz).  The GPU test runs the `warpdemux predict <prep_dir>` flow end to end."""
import gzip
import io
import json
import os
from types import SimpleNamespace

import numpy as np
import pandas as pd
import pytest

from conftest import GOLD


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(GOLD, "io_formats.npz")) as z:
        return {k: z[k] for k in z.files}


def _write_prep_dir(gold, root):
    from warpdemux_b200.io import results as R

    os.makedirs(os.path.join(root, "fingerprints"), exist_ok=True)
    lo = 0
    for bidx, cnt in enumerate(gold["n_per_file"]):
        sl = slice(lo, lo + int(cnt))
        lo += int(cnt)
        res = [SimpleNamespace(read_id=str(r), barcode_fpt=f, dwell_times=d)
               for r, f, d in zip(gold["read_ids"][sl], gold["fpt"][sl], gold["dwell"][sl])]
        R.save_fpts_signals(res, os.path.join(root, "fingerprints", f"barcode_fpts_{bidx}.npz"),
                            save_dwell_time=(bidx == 0))
    return [os.path.join(root, "fingerprints", f"barcode_fpts_{b}.npz") for b in range(len(gold["n_per_file"]))]


def test_fingerprint_npz_matches_reference(gold, tmp_path):
    files = _write_prep_dir(gold, str(tmp_path))
    for bidx, fn in enumerate(files):
        with np.load(fn) as w:
            assert list(w.files) == json.loads(str(gold[f"fpts{bidx}_keys"]))
            for k in w.files:
                want = gold[f"fpts{bidx}_{k}"]
                assert w[k].dtype == want.dtype and w[k].shape == want.shape, (bidx, k)
                assert np.array_equal(w[k], want), (bidx, k)


def test_fingerprint_npz_from_arrays(gold, tmp_path):
    from warpdemux_b200.io import results as R

    n0 = int(gold["n_per_file"][0])
    fn = str(tmp_path / "barcode_fpts_0.npz")
    R.save_fpts_arrays(gold["read_ids"][:n0], gold["fpt"][:n0], fn, dwell_times=gold["dwell"][:n0])
    with np.load(fn) as w:
        for k in ("num_reads", "read_ids", "signals", "dwell_times"):
            assert np.array_equal(w[k], gold[f"fpts0_{k}"]) and w[k].dtype == gold[f"fpts0_{k}"].dtype


def test_yield_fpts_from_npz_matches_reference(gold, tmp_path):
    from warpdemux_b200.io import results as R

    files = _write_prep_dir(gold, str(tmp_path))
    excl, incl = set(gold["excl"].tolist()), set(gold["incl"].tolist())
    for tag, kw in (("all", dict(read_ids_incl=set(), read_ids_excl=set())),
                    ("excl", dict(read_ids_incl=set(), read_ids_excl=excl)),
                    ("incl", dict(read_ids_incl=incl, read_ids_excl=set()))):
        batches = list(R.yield_fpts_from_npz(files, batch_size=32, **kw))
        assert [len(b[1]) for b in batches] == gold[f"yield_{tag}_sizes"].tolist(), tag
        assert np.array_equal(np.concatenate([b[1] for b in batches]).astype(str), gold[f"yield_{tag}_ids"]), tag
        assert np.array_equal(np.concatenate([b[0] for b in batches], axis=0), gold[f"yield_{tag}_fpts"]), tag
    # a file without any listed read (the reference raises IndexError there) is simply skipped
    only_first = set(gold["read_ids"][:3].tolist())
    got = list(R.yield_fpts_from_npz(files, read_ids_incl=only_first, read_ids_excl=set(), batch_size=32))
    assert len(got) == 1 and set(got[0][1].tolist()) == only_first
    # include and exclude together: exclude wins (file_proc.py:288-290)
    got = list(R.yield_fpts_from_npz(files, read_ids_incl=only_first, read_ids_excl={gold["read_ids"][0]}, batch_size=32))
    assert sorted(got[0][1].tolist()) == sorted(gold["read_ids"][1:3].tolist())


def _ref_frames(gold):
    return [pd.read_csv(io.StringIO(str(gold[f"pred{b}_csv"]))) for b in range(len(gold["n_per_file"]))]


def test_prediction_csv_and_resume_scan_match_reference(gold, tmp_path):
    """Re-serialising the reference's own DataFrame through our writer gives the reference's bytes, and the
    resume scan sees the same read ids and batch indices."""
    from warpdemux_b200.io import results as R

    root = str(tmp_path)
    os.makedirs(os.path.join(root, "predictions"))
    _write_prep_dir(gold, root)
    for bidx, df in enumerate(_ref_frames(gold)):
        ids = df.pop("#read_id")
        df2 = R.add_read_id_col_to_predictions(df, ids.to_numpy())
        assert df2.columns[0] == "#read_id"
        with pytest.raises(ValueError):
            R.add_read_id_col_to_predictions(df2, ids.to_numpy())
        fn = os.path.join(root, "predictions", f"barcode_predictions_{bidx}.csv.gz")
        R.save_predictions(df2, fn)
        with gzip.open(fn, "rt") as fh:
            assert fh.read() == str(gold[f"pred{bidx}_csv"])
    ids, mp, mf = R.scan_processed_reads(root, scan_failed=False, result_type="predictions")
    assert sorted(ids) == gold["scan_pred_ids"].tolist() and [mp, mf] == gold["scan_pred_bidx"].tolist()
    ids, mp, mf = R.scan_processed_reads(root, scan_failed=False, result_type="fingerprints")
    assert sorted(str(r) for r in ids) == gold["scan_fpts_ids"].tolist() and [mp, mf] == gold["scan_fpts_bidx"].tolist()
    with pytest.raises(ValueError):
        R.scan_processed_reads(root, result_type="boundaries")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["exact", "guarded"])
def test_predict_fingerprint_dir_matches_reference_csv(gold, tmp_path, mode):
    """`warpdemux predict <prep_dir>` on the GPU: the csv.gz files hold the reference's text, re-cut into
    output batches like _queue_batch_processor_df, and a resumed run only adds the missing reads."""
    from warpdemux_b200.io import results as R
    from warpdemux_b200.models.dtw_svm import DTW_SVM

    prep, out = str(tmp_path / "prep"), str(tmp_path / "out")
    _write_prep_dir(gold, prep)
    model = DTW_SVM.load(os.path.join(GOLD, "models", "WDX4_rna004_v1_0.npz"), mode=mode)
    n, nf = R.predict_fingerprint_dir(model, prep, out, minibatch_size=64, batch_size_output=50)
    assert (n, nf) == (120, 3)
    got = pd.concat([pd.read_csv(os.path.join(out, "predictions", f"barcode_predictions_{b}.csv.gz")) for b in range(3)],
                    ignore_index=True)
    want = pd.concat(_ref_frames(gold), ignore_index=True)
    assert list(got.columns) == list(want.columns)
    assert got["#read_id"].tolist() == want["#read_id"].tolist()
    assert got["predicted_barcode"].tolist() == want["predicted_barcode"].tolist()
    # rounded to 3 / 4 decimals from probabilities that agree to 2e-6: at most one unit in the last place
    assert np.abs(got["confidence_score"] - want["confidence_score"]).max() <= 1e-3 + 1e-12
    pc = [c for c in want.columns if c.startswith("p") and c != "predicted_barcode"]
    assert np.abs(got[pc].to_numpy() - want[pc].to_numpy()).max() <= 1e-4 + 1e-12
    assert (got[pc].to_numpy() != want[pc].to_numpy()).mean() < 0.01
    with gzip.open(os.path.join(out, "predictions", "barcode_predictions_0.csv.gz"), "rt") as fh:
        assert fh.readline().strip() == str(gold["pred0_csv"]).splitlines()[0]
    # resume: drop the last file, continue from the output directory
    os.remove(os.path.join(out, "predictions", "barcode_predictions_2.csv.gz"))
    n2, nf2 = R.predict_fingerprint_dir(model, prep, out, minibatch_size=64, batch_size_output=50, continue_from=out)
    assert (n2, nf2) == (20, 1)
    again = pd.read_csv(os.path.join(out, "predictions", "barcode_predictions_2.csv.gz"))
    assert again["#read_id"].tolist() == want["#read_id"].tolist()[100:]
    model.close() if hasattr(model, "close") else None


# ---- detected_boundaries_*.csv.gz / failed_reads_*.csv.gz (file_proc.py:650-724, adapted/output.py:26-51) -------------
@pytest.fixture(scope="module")
def gold_bounds():
    with np.load(os.path.join(GOLD, "boundaries_rna004.npz")) as z:
        return {k: z[k] for k in z.files}


def _unjson(v):
    if isinstance(v, dict):
        if "__nd__" in v:
            return np.array(v["__nd__"], dtype=v["dtype"])
        if "__np_int__" in v:
            return np.int64(v["__np_int__"])
        return np.dtype(v["dtype"]).type(v["__np_float__"])
    return v


def _mirror_results(records):
    """The reference's `to_summary_dict()` records -> this package's ReadResult / DetectResults mirrors."""
    from warpdemux_b200.detect.combined import DetectResults
    from warpdemux_b200.sig_proc import ReadResult

    rr_fields = ("adapter_dt_med", "adapter_dt_mad", "adapter_event_mean", "adapter_event_std", "adapter_event_med",
                 "adapter_event_mad", "seg_cons_query_start", "seg_cons_query_end", "sig_barcode_start")
    out = []
    for rec in records:
        rec = {k: _unjson(v) for k, v in rec.items()}
        dr = DetectResults(success=bool(rec["success"]))
        for k, v in rec.items():
            if k not in rr_fields and k not in ("read_id", "fail_reason", "success"):
                assert hasattr(dr, k), k
                setattr(dr, k, v)
        out.append(ReadResult(read_id=rec["read_id"], success=bool(rec["success"]) and rec["fail_reason"] == "", fail_reason=rec["fail_reason"],
                              detect_results=dr, **{k: rec[k] for k in rr_fields}))
    return out


def test_boundary_tables_match_reference_bytes(gold_bounds, tmp_path):
    """The writers give the reference's bytes for the reference's own records (columns, order, None / NaN cells, array
    cells, 3-decimal rounding), and RunDirWriter re-cuts into output batches / resumes like the reference's collectors."""
    from warpdemux_b200.io import results as R

    g = gold_bounds
    passed = _mirror_results(json.loads(str(g["records_pass"])))
    failed = _mirror_results(json.loads(str(g["records_fail"])))
    d = str(tmp_path)
    R.save_detect_results("pass", passed, 0, save_fpts=False, output_dir_boundaries=d, output_dir_fpts=d)
    R.save_detect_results("fail", failed, 0, output_dir_fail=d, save_fpts=False, save_dwell_time=False, save_boundaries=True)
    assert gzip.open(os.path.join(d, "detected_boundaries_0.csv.gz"), "rb").read() == bytes(g["csv_pass"])
    assert gzip.open(os.path.join(d, "failed_reads_0.csv.gz"), "rb").read() == bytes(g["csv_fail"])
    with pytest.raises(ValueError):
        R.save_detect_results("maybe", passed, 0)
    # collectors: 100 rows per file, remainder at close; a resumed run continues the numbering and knows the processed reads
    run = str(tmp_path / "run")
    w = R.RunDirWriter(run, batch_size_output=100, save_predictions=False, save_fpts=False)
    w.add(passed[:150] + failed[:30])
    w.add(passed[150:] + failed[30:])
    w.close()
    names = sorted(os.path.basename(f) for f in w.files_written)
    n_pass, n_fail = len(passed), len(failed)
    assert names == sorted([f"detected_boundaries_{i}.csv.gz" for i in range(-(-n_pass // 100))] + [f"failed_reads_{i}.csv.gz" for i in range(-(-n_fail // 100))])
    body = b"".join(gzip.open(os.path.join(run, "boundaries", f"detected_boundaries_{i}.csv.gz"), "rb").read().split(b"\n", 1)[1]
                    for i in range(-(-n_pass // 100)))
    assert body.count(b"\n") == n_pass
    os.makedirs(os.path.join(run, "fingerprints"), exist_ok=True)
    R.save_fpts_arrays(np.array([r.read_id for r in passed]), np.zeros((n_pass, 25)), os.path.join(run, "fingerprints", "barcode_fpts_2.npz"))
    w2 = R.RunDirWriter(str(tmp_path / "run2"), batch_size_output=100, save_predictions=False, save_fpts=True, continue_from=run)
    assert w2.processed == {r.read_id for r in passed + failed}
    assert w2.bidx == {"pass": 3, "fail": -(-n_fail // 100), "predict": 3}


@pytest.mark.gpu
def test_gpu_run_directory_equals_reference_tables(gold_bounds, tmp_path, models):
    """Raw minibatch rows -> CNN -> validation / LLR fallback -> fingerprints -> calls on the GPU, written through
    RunDirWriter: detected_boundaries_0.csv.gz and failed_reads_0.csv.gz are byte-identical to the files the reference wrote
    for the same minibatch (439 real reads: every failed read, every LLR / hail-mary read, every 16th read)."""
    import test_gpu_llr as T
    from wdx_testutil import real4000_rows

    from warpdemux_b200.detect import cnn, combined
    from warpdemux_b200.file_proc import demux_minibatches_to_dir
    from warpdemux_b200.io import results as R
    from warpdemux_b200.models.dtw_svm import DTW_SVM
    from warpdemux_b200.sig_proc import FingerprintConfig

    with np.load(os.path.join(GOLD, "real4000_rna004_WDX4.npz")) as z:
        g = {k: z[k] for k in z.files}
    idx, rows, _, _ = real4000_rows(g)
    assert np.array_equal(idx, gold_bounds["subset"])
    spc = T._spc(g)
    spc.primary_method = "cnn"
    md = cnn.load_cnn_model(os.path.join(GOLD, "models", "cnn_rna004_130bps_v0.2.4.npz"), device=0)
    mp = DTW_SVM(models["WDX4_rna004_v1_0"], device=0, mode="exact")
    v = combined.Validator(combined.ValidateConfig(), device=0, llr=T._llr_cfg(g))
    run = str(tmp_path / "run")
    w = R.RunDirWriter(run, batch_size_output=4000, save_fpts=True, save_dwell_time=True)
    counts = demux_minibatches_to_dir([(rows, g["full_lengths"][idx], [str(x) for x in g["read_ids"][idx]])], mp, md, spc, w,
                                      validator=v, fp_config=FingerprintConfig(**T._fp_cfg(g)))
    v.close()
    md.close()
    assert counts == {"reads": idx.size, "pass": len(gold_bounds["pass_ids"]), "fail": len(gold_bounds["fail_ids"]), "skipped": 0}

    def first_diff(got, want):
        gl, wl = got.split(b"\n"), want.split(b"\n")
        for a, b in zip(gl, wl):
            if a != b:
                ga, wb = a.split(b","), b.split(b",")
                cols = wl[0].split(b",")
                return [(cols[i] if i < len(cols) else i, x, y) for i, (x, y) in enumerate(zip(ga, wb)) if x != y][:6], a[:36]
        return len(gl), len(wl)

    got_pass = gzip.open(os.path.join(run, "boundaries", "detected_boundaries_0.csv.gz"), "rb").read()
    got_fail = gzip.open(os.path.join(run, "failed_reads", "failed_reads_0.csv.gz"), "rb").read()
    # the one read whose change points hang on how numpy's unstable argsort orders two EQUAL t-test scores (fixture key
    # stable_idx, see test_gpu_llr.py): its six adapter-event statistics follow the kernels' stable order
    ties = [str(g["read_ids"][i]).encode() for i in g["stable_idx"]]
    assert len(ties) <= 2
    drop = lambda blob: b"\n".join(ln for ln in blob.split(b"\n") if not any(ln.startswith(t) for t in ties))
    assert drop(got_pass) == drop(bytes(gold_bounds["csv_pass"])), first_diff(drop(got_pass), drop(bytes(gold_bounds["csv_pass"])))
    assert got_pass.count(b"\n") == bytes(gold_bounds["csv_pass"]).count(b"\n")
    assert got_fail == bytes(gold_bounds["csv_fail"]), first_diff(got_fail, bytes(gold_bounds["csv_fail"]))
    # the run directory resumes: everything is already processed
    ids, mp_, mf_ = R.scan_processed_reads(run, scan_failed=True, result_type="predictions")
    assert len(ids) == idx.size and (mp_, mf_) == (0, 0)
    with np.load(os.path.join(run, "fingerprints", "barcode_fpts_0.npz")) as z:
        assert list(z["read_ids"]) == [str(x) for x in gold_bounds["pass_ids"]]

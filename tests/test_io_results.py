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

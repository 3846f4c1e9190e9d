"""Result formats either side of the hot path — host-side mirrors of the
reference's functions of the same names in warpdemux/file_proc.py, so that the GPU path reads what
`warpdemux prep` wrote and writes what downstream tools read:

    fingerprints/barcode_fpts_{i}.npz          save_fpts_signals         file_proc.py:725-754
        num_reads, read_ids, signals[, dwell_times]
    predictions/barcode_predictions_{i}.csv.gz save_predictions          file_proc.py:757-766
        header  #read_id,predicted_barcode,confidence_score,pXX...,p-1   (models/utils.py:36-43)
    add_read_id_col_to_predictions                                       file_proc.py:769-780
    boundaries/detected_boundaries_{i}.csv.gz  save_detect_results("pass") file_proc.py:682-724 -> adapted/output.py:26-51
    failed_reads/failed_reads_{i}.csv.gz       save_batch_outputs_fail     file_proc.py:650-665
        one row per read: read_id, the DetectResults fields (container_types.py:14-77 minus success / llr_trace),
        [fail_reason,] the fingerprint stage's adapter statistics (sig_proc.py:44-59); floats rounded to 3 decimals
    scan_processed_reads (resume)                                        file_proc.py:129-169
    yield_fpts_from_npz (input of `warpdemux predict`)                   file_proc.py:282-330

`predict_fingerprint_dir` is the `warpdemux predict <prep_dir>` call stack (SURVEY 3.2:
worker_enqueue_minibatches_fpts -> worker_predict_on_preloaded_fpts -> queue_batch_processor ->
save_batch_predictions, file_proc.py:357-377, 457-497, 500-541, 667-679) with the process pool replaced by
one GPU-owning process: minibatches are predicted by `DTW_SVM.predict` (one C-ABI call each) and re-cut
into output files of `batch_size_output` rows exactly like `_queue_batch_processor_df`.

Deviations from the reference, both in the direction of not crashing:
  * `yield_fpts_from_npz`: a file that holds none of the listed read ids is skipped / passed through; the
    reference indexes with an empty float64 array there and raises IndexError (file_proc.py:303-312).
  * `scan_processed_reads(scan_failed=True)`: failed_reads_*.csv.gz is opened with gzip; the reference opens
    the gzip file as plain text (file_proc.py:147).
"""
from __future__ import annotations

import gzip
import os
from typing import Generator, Iterable, List, Optional, Sequence, Set, Tuple, Union

import numpy as np
import pandas as pd


def save_fpts_signals(list_of_processing_results: Sequence, filename: str, save_dwell_time: bool = True):
    """`ReadResult`s (anything with read_id / barcode_fpt / dwell_times) -> npz; returns what the reference returns."""
    read_ids = np.array([res.read_id for res in list_of_processing_results])
    barcode_fpts = np.array([res.barcode_fpt for res in list_of_processing_results])
    dwell_times = np.array([res.dwell_times for res in list_of_processing_results])
    num_reads = len(read_ids)
    if save_dwell_time:
        np.savez(filename, num_reads=num_reads, read_ids=read_ids, signals=barcode_fpts, dwell_times=dwell_times)
    else:
        np.savez(filename, num_reads=num_reads, read_ids=read_ids, signals=barcode_fpts)
    return read_ids, barcode_fpts, dwell_times


def save_fpts_arrays(read_ids: np.ndarray, barcode_fpts: np.ndarray, filename: str,
                     dwell_times: Optional[np.ndarray] = None) -> None:
    """Same file from the arrays a `FingerprintBatch` already holds (no per-read Python objects)."""
    kw = dict(num_reads=len(read_ids), read_ids=np.asarray(read_ids), signals=np.asarray(barcode_fpts))
    if dwell_times is not None:
        kw["dwell_times"] = np.asarray(dwell_times)
    np.savez(filename, **kw)


def save_detected_boundaries(processing_results: Sequence, filename: str, save_fail_reasons: bool = False) -> None:
    """adapted/output.py:26-51: `ReadResult.to_summary_dict()` rows -> csv(.gz), floats rounded to 3 decimals.  The table's
    columns, their order and the text of every cell are fixed by the reference's dataclasses and by pandas."""
    df = pd.DataFrame([pr.to_summary_dict() for pr in processing_results])
    if not df.empty:
        drop = ["success", "llr_trace"] + ([] if save_fail_reasons else ["fail_reason"])
        df = df.drop(columns=[c for c in drop if c in df.columns])
    df.round(3).to_csv(filename, index=False)


def save_detect_results(pass_or_fail: str, results: Sequence, batch_idx: int, save_fpts: bool = True, save_dwell_time: bool = False,
                        save_boundaries: bool = True, output_dir_boundaries: str = "", output_dir_fpts: str = "",
                        output_dir_fail: str = "", **kwargs):
    """file_proc.py:682-724: detected_boundaries_{i}.csv.gz (+ barcode_fpts_{i}.npz) for passed reads,
    failed_reads_{i}.csv.gz (with the fail reasons) for failed ones."""
    if pass_or_fail == "pass":
        fn, dirn = "detected_boundaries", output_dir_boundaries
    elif pass_or_fail == "fail":
        fn, dirn = "failed_reads", output_dir_fail
    else:
        raise ValueError(f"Invalid pass_or_fail: {pass_or_fail}. Must be 'pass' or 'fail'.")
    if save_boundaries:
        save_detected_boundaries(results, os.path.join(dirn, f"{fn}_{batch_idx}.csv.gz"), save_fail_reasons=pass_or_fail == "fail")
    if pass_or_fail == "pass" and save_fpts:
        read_ids, signals, _ = save_fpts_signals(results, os.path.join(output_dir_fpts, f"barcode_fpts_{batch_idx}.npz"),
                                                 save_dwell_time=save_dwell_time)
        return read_ids, signals
    return None


class RunDirWriter:
    """The three output collectors of the reference's demux / prep run (file_proc.py:500-584 `queue_batch_processor` with
    save_batch_outputs_pass / _fail / save_batch_predictions, :627-679) without the queues: results are re-cut into files
    of `batch_size_output` rows, batch indices continue after the ones `continue_from` already holds
    (handle_previous_results, file_proc.py:172-185), directory names are OutputConfig's (config/file_proc.py:18-49)."""

    def __init__(self, output_dir: str, batch_size_output: int = 4000, save_predictions: bool = True, save_fpts: bool = False,
                 save_dwell_time: bool = False, save_boundaries: bool = True, continue_from: Optional[str] = None):
        self.dir_pred = os.path.join(output_dir, "predictions")
        self.dir_fail = os.path.join(output_dir, "failed_reads")
        self.dir_fpts = os.path.join(output_dir, "fingerprints")
        self.dir_boundaries = os.path.join(output_dir, "boundaries")
        self.batch = int(batch_size_output)
        self.opt = dict(save_predictions=save_predictions, save_fpts=save_fpts, save_dwell_time=save_dwell_time,
                        save_boundaries=save_boundaries)
        os.makedirs(self.dir_fail, exist_ok=True)
        for on, d in ((save_predictions, self.dir_pred), (save_boundaries, self.dir_boundaries), (save_fpts, self.dir_fpts)):
            if on:
                os.makedirs(d, exist_ok=True)
        self.processed: Set[str] = set()
        self.bidx = {"pass": 0, "fail": 0, "predict": 0}
        if continue_from:
            result_type = "predictions" if save_predictions else "fingerprints"
            self.processed, max_pass, max_fail = scan_processed_reads(continue_from, scan_failed=True, result_type=result_type)
            self.bidx = {"pass": max_pass + 1, "fail": max_fail + 1, "predict": max_pass + 1}
        self._pass: List = []
        self._fail: List = []
        self._pred: List[pd.DataFrame] = []
        self._pred_rows = 0
        self.files_written: List[str] = []

    def _flush_pass(self, rows: Sequence) -> None:
        i = self.bidx["pass"]
        if self.opt["save_boundaries"] or self.opt["save_fpts"]:
            save_detect_results("pass", rows, i, save_fpts=self.opt["save_fpts"], save_dwell_time=self.opt["save_dwell_time"],
                                save_boundaries=self.opt["save_boundaries"], output_dir_boundaries=self.dir_boundaries,
                                output_dir_fpts=self.dir_fpts)
            if self.opt["save_boundaries"]:
                self.files_written.append(os.path.join(self.dir_boundaries, f"detected_boundaries_{i}.csv.gz"))
            if self.opt["save_fpts"]:
                self.files_written.append(os.path.join(self.dir_fpts, f"barcode_fpts_{i}.npz"))
        self.bidx["pass"] += 1

    def _flush_fail(self, rows: Sequence) -> None:
        i = self.bidx["fail"]
        save_detect_results("fail", rows, i, output_dir_fail=self.dir_fail, save_fpts=False, save_dwell_time=False, save_boundaries=True)
        self.files_written.append(os.path.join(self.dir_fail, f"failed_reads_{i}.csv.gz"))
        self.bidx["fail"] += 1

    def _flush_pred(self, df: pd.DataFrame) -> None:
        i = self.bidx["predict"]
        path = os.path.join(self.dir_pred, f"barcode_predictions_{i}.csv.gz")
        save_predictions(df, path)
        self.files_written.append(path)
        self.bidx["predict"] += 1

    def add(self, read_results: Sequence, predictions: Optional[pd.DataFrame] = None) -> None:
        """One minibatch: `ReadResult`s (split here into pass / fail like file_proc.py:418-440) and the predictions table of
        its passed reads ('#read_id' column first)."""
        for r in read_results:
            (self._pass if r.success else self._fail).append(r)
        while len(self._pass) >= self.batch:                       # _queue_batch_processor_list
            self._flush_pass(self._pass[:self.batch])
            self._pass = self._pass[self.batch:]
        while len(self._fail) >= self.batch:
            self._flush_fail(self._fail[:self.batch])
            self._fail = self._fail[self.batch:]
        if predictions is not None and self.opt["save_predictions"] and len(predictions):
            self._pred.append(predictions)
            self._pred_rows += len(predictions)
            while self._pred_rows >= self.batch:                   # _queue_batch_processor_df
                cur = pd.concat(self._pred, axis=0)
                self._flush_pred(cur.iloc[:self.batch].copy())
                self._pred = [cur.iloc[self.batch:]]
                self._pred_rows -= self.batch

    def close(self) -> None:
        """The remainders (file_proc.py:577-584)."""
        if self._pass:
            self._flush_pass(self._pass)
        if self._fail:
            self._flush_fail(self._fail)
        if self._pred_rows > 0:
            self._flush_pred(pd.concat(self._pred, axis=0))
        self._pass, self._fail, self._pred, self._pred_rows = [], [], [], 0


def save_predictions(predictions: pd.DataFrame, filename: str) -> None:
    predictions.to_csv(filename, index=False, compression="gzip")


def add_read_id_col_to_predictions(predictions: pd.DataFrame, read_ids: Union[List[str], np.ndarray]) -> pd.DataFrame:
    cols = predictions.columns.tolist()
    if "#read_id" in cols:
        raise ValueError("'#read_id' already in dataframe")
    predictions["#read_id"] = read_ids
    return predictions[["#read_id", *cols]]


def determine_bidx_from_file(file: str) -> int:
    """file_proc.py:119-120."""
    return int(file.split("_")[-1].split(".")[0])


def scan_processed_reads(continue_from_path: str, scan_failed: bool = False,
                         result_type: str = "predictions") -> Tuple[Set[str], int, int]:
    processed_reads: Set[str] = set()
    max_pass_bidx = -1
    max_fail_bidx = -1
    if result_type not in ["predictions", "fingerprints"]:
        raise ValueError(f"Invalid result_type: {result_type}. Must be 'predictions' or 'fingerprints'.")
    if scan_failed:
        fail_sub = os.path.join(continue_from_path, "failed_reads")
        for file in os.listdir(fail_sub):
            if file.startswith("failed_reads_") and file.endswith(".csv.gz"):
                max_fail_bidx = max(max_fail_bidx, determine_bidx_from_file(file))
                with gzip.open(os.path.join(fail_sub, file), "rt") as f:
                    processed_reads.update(line.split(",")[0] for line in f.readlines()[1:])
    if result_type == "predictions":
        pass_sub, starts_with, extension = os.path.join(continue_from_path, "predictions"), "barcode_predictions_", "csv.gz"
    else:
        pass_sub, starts_with, extension = os.path.join(continue_from_path, "fingerprints"), "barcode_fpts_", "npz"
    for file in os.listdir(pass_sub):
        if file.startswith(starts_with) and file.endswith(extension):
            max_pass_bidx = max(max_pass_bidx, determine_bidx_from_file(file))
            if extension == "csv.gz":
                with gzip.open(os.path.join(pass_sub, file), "rt") as f:
                    processed_reads.update(line.split(",")[0] for line in f.readlines()[1:])
            else:
                with np.load(os.path.join(pass_sub, file)) as npz:
                    processed_reads.update(npz["read_ids"])
    return processed_reads, max_pass_bidx, max_fail_bidx


def yield_fpts_from_npz(npz_files: Iterable[str], read_ids_incl: Set[str], read_ids_excl: Set[str],
                        batch_size: int) -> Generator[Tuple[np.ndarray, np.ndarray], None, None]:
    """Minibatches of exactly `batch_size` fingerprints (the last one shorter) across the given files,
    in file order, honouring the include / exclude sets like the reference."""
    if read_ids_incl and read_ids_excl:
        read_ids_incl = read_ids_incl.difference(read_ids_excl)
        read_ids_excl = set()
    N = batch_size
    fpts = np.empty((0, 0), dtype=np.float32)
    read_ids = np.empty(0, dtype=object)
    for filename in npz_files:
        with np.load(filename) as data:
            file_fpts = data["signals"]
            file_read_ids = data["read_ids"]
        if read_ids_excl:
            keep = np.array([rid not in read_ids_excl for rid in file_read_ids], dtype=bool)
            file_fpts, file_read_ids = file_fpts[keep], file_read_ids[keep]
        elif read_ids_incl:
            keep = np.array([rid in read_ids_incl for rid in file_read_ids], dtype=bool)
            file_fpts, file_read_ids = file_fpts[keep], file_read_ids[keep]
        if fpts.size > 0:
            fpts = np.concatenate((fpts, file_fpts), axis=0)
            read_ids = np.concatenate((read_ids, file_read_ids), axis=0)
        else:
            fpts, read_ids = file_fpts, file_read_ids
        while len(fpts) >= N:
            yield fpts[:N].copy(), read_ids[:N].copy()
            fpts, read_ids = fpts[N:], read_ids[N:]
    if fpts.size > 0:
        yield fpts, read_ids


def list_fingerprint_files(prep_dir: str) -> List[str]:
    """`<prep_dir>/fingerprints/barcode_fpts_*.npz` in batch-index order (parser.py:435-441 globs the same pattern)."""
    sub = os.path.join(prep_dir, "fingerprints")
    files = [f for f in os.listdir(sub) if f.startswith("barcode_fpts_") and f.endswith(".npz")]
    return [os.path.join(sub, f) for f in sorted(files, key=determine_bidx_from_file)]


def predict_fingerprint_dir(model, prep_dir: str, output_dir: str, minibatch_size: int = 1 << 20,
                            batch_size_output: int = 4000, continue_from: Optional[str] = None,
                            read_ids_incl: Optional[Set[str]] = None) -> Tuple[int, int]:
    """`warpdemux predict`: fingerprints on disk -> predictions/barcode_predictions_{i}.csv.gz under
    `output_dir`.  `model` is a `DTW_SVM` (its `predict(..., return_df=True)` is the reference seam);
    `continue_from` resumes like `handle_previous_results` (file_proc.py:172-185): reads already in its
    predictions are skipped and the batch index continues after the highest one found.
    Returns (reads predicted, files written).  `batch_size_output` defaults to the reference's 4000
    (config `batch.batch_size_output`)."""
    excl: Set[str] = set()
    bidx = 0
    if continue_from:
        excl, max_pass, _ = scan_processed_reads(continue_from, scan_failed=False, result_type="predictions")
        bidx = max_pass + 1
    out_sub = os.path.join(output_dir, "predictions")
    os.makedirs(out_sub, exist_ok=True)
    pending: List[pd.DataFrame] = []
    pending_rows = 0
    n_reads = n_files = 0

    def flush(df: pd.DataFrame) -> None:
        nonlocal bidx, n_files
        save_predictions(df, os.path.join(out_sub, f"barcode_predictions_{bidx}.csv.gz"))
        bidx += 1
        n_files += 1

    for fpts, read_ids in yield_fpts_from_npz(list_fingerprint_files(prep_dir), read_ids_incl or set(), excl,
                                              minibatch_size):
        df = model.predict(fpts, pbar=False, nproc=1, return_df=True)          # file_proc.py:488-493
        df = add_read_id_col_to_predictions(df, read_ids)
        n_reads += len(df)
        pending.append(df)
        pending_rows += len(df)
        while pending_rows >= batch_size_output:                                # _queue_batch_processor_df
            cur = pd.concat(pending, axis=0)
            flush(cur.iloc[:batch_size_output].copy())
            pending = [cur.iloc[batch_size_output:]]
            pending_rows -= batch_size_output
    if pending_rows > 0:
        flush(pd.concat(pending, axis=0))
    return n_reads, n_files

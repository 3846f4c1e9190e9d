"""Result formats either side of the hot path — host-side mirrors of the
reference's functions of the same names in warpdemux/file_proc.py, so that the GPU path reads what
`warpdemux prep` wrote and writes what downstream tools read:

    fingerprints/barcode_fpts_{i}.npz          save_fpts_signals         file_proc.py:725-754
        num_reads, read_ids, signals[, dwell_times]
    predictions/barcode_predictions_{i}.csv.gz save_predictions          file_proc.py:757-766
        header  #read_id,predicted_barcode,confidence_score,pXX...,p-1   (models/utils.py:36-43)
    add_read_id_col_to_predictions                                       file_proc.py:769-780
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

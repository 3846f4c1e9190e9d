"""Minimal pod5 reader (host side) — enough to feed the classification path from
the reference's own test file without the `pod5` package.

Replaces, for the fields the path needs, what `warpdemux/file_proc.py:227-279`
gets from `pod5.Reader`: `read_id`, `num_samples` and `signal_pa`
(= (adc + calibration_offset) * calibration_scale, float32).

Format (pod5 v0.2.x/0.3.x container): a signature, then three embedded Arrow IPC
files (signal table, run-info table, reads table), each delimited by `ARROW1`
magic strings.  Signal rows are VBZ chunks: zstd( streamvbyte16( zigzag( delta(
int16 samples )))), with one key BIT per sample (0 = 1 byte, 1 = 2 bytes, LSB
first) ahead of the data bytes.  A read's signal is the concatenation of the
signal-table rows listed in the reads table.

Parity: there is no `pod5` package in this image to compare with; the decoder is
self-checked (decoded length == `samples`, consumed bytes == chunk length) in
tests/test_pod5_reader.py — unpinned by the reference.
"""
from __future__ import annotations

import re
import uuid
from dataclasses import dataclass
from typing import Iterator, List, Optional, Sequence

import numpy as np

_SIGNATURE = b"\x8bPOD\r\n\x1a\n"


def decode_vbz(chunk: bytes, n_samples: int) -> np.ndarray:
    """One VBZ signal chunk -> int16[n_samples]."""
    import pyarrow as pa

    raw = pa.CompressedInputStream(pa.BufferReader(chunk), "zstd").read()
    buf = np.frombuffer(raw, dtype=np.uint8)
    if n_samples == 0:
        return np.zeros(0, dtype=np.int16)
    nk = (n_samples + 7) // 8
    keys = np.unpackbits(buf[:nk], bitorder="little")[:n_samples].astype(np.int64)
    width = 1 + keys
    off = np.cumsum(width) - width
    data = buf[nk:]
    if off[-1] + width[-1] != data.size:
        raise ValueError("VBZ chunk length does not match its key stream")
    lo = data[off].astype(np.uint32)
    hi = data[np.minimum(off + 1, data.size - 1)].astype(np.uint32) * keys.astype(np.uint32)
    u = lo | (hi << 8)
    zz = (u >> 1).astype(np.int32) ^ -(u & 1).astype(np.int32)      # zig-zag
    return np.cumsum(zz, dtype=np.int32).astype(np.int16)            # delta (wraps like int16 arithmetic)


@dataclass
class Pod5Read:
    read_id: str
    num_samples: int
    calibration_offset: float
    calibration_scale: float
    signal_rows: List[int]
    _file: "Pod5File"

    @property
    def signal(self) -> np.ndarray:
        """Raw ADC samples, int16."""
        return self._file._signal(self.signal_rows)

    @property
    def signal_pa(self) -> np.ndarray:
        """Calibrated picoampere signal, float32 (pod5 `ReadRecord.signal_pa`)."""
        return ((self.signal.astype(np.float32) + np.float32(self.calibration_offset))
                * np.float32(self.calibration_scale))


class Pod5File:
    def __init__(self, path: str):
        import pyarrow as pa

        self.path = path
        with open(path, "rb") as fh:
            data = fh.read()
        if not (data.startswith(_SIGNATURE) and data.endswith(_SIGNATURE)):
            raise ValueError(f"{path}: not a pod5 file")
        marks = [m.start() for m in re.finditer(b"ARROW1", data)]
        if len(marks) < 6 or len(marks) % 2:
            raise ValueError(f"{path}: cannot locate the embedded Arrow tables")
        tables = {}
        for a, b in zip(marks[0::2], marks[1::2]):
            rd = pa.ipc.open_file(pa.BufferReader(data[a:b + 6]))
            names = set(rd.schema.names)
            if {"signal", "samples"} <= names:
                tables["signal"] = rd
            elif {"num_samples", "calibration_scale"} <= names:
                tables["reads"] = rd
        if set(tables) != {"signal", "reads"}:
            raise ValueError(f"{path}: signal / reads tables not found")
        self._sig = tables["signal"]
        self._reads = tables["reads"].read_all()
        # row index -> (batch, offset in batch) for the signal table
        counts = [self._sig.get_batch(i).num_rows for i in range(self._sig.num_record_batches)]
        self._batch_start = np.concatenate([[0], np.cumsum(counts)])
        self._batch_cache = {}

    def __len__(self) -> int:
        return self._reads.num_rows

    def _signal_row(self, row: int) -> np.ndarray:
        b = int(np.searchsorted(self._batch_start, row, side="right") - 1)
        if b not in self._batch_cache:
            if len(self._batch_cache) > 4:
                self._batch_cache.clear()
            self._batch_cache[b] = self._sig.get_batch(b)
        batch = self._batch_cache[b]
        i = row - int(self._batch_start[b])
        return decode_vbz(batch.column("signal")[i].as_py(), int(batch.column("samples")[i].as_py()))

    def _signal(self, rows: Sequence[int]) -> np.ndarray:
        parts = [self._signal_row(int(r)) for r in rows]
        return parts[0] if len(parts) == 1 else np.concatenate(parts)

    def reads(self, selection: Optional[Sequence[str]] = None) -> Iterator[Pod5Read]:
        t = self._reads
        ids = t.column("read_id").to_pylist()
        ns = t.column("num_samples").to_pylist()
        off = t.column("calibration_offset").to_pylist()
        sc = t.column("calibration_scale").to_pylist()
        rows = t.column("signal").to_pylist()
        want = set(selection) if selection is not None else None
        for i in range(t.num_rows):
            rid = str(uuid.UUID(bytes=ids[i]))
            if want is not None and rid not in want:
                continue
            yield Pod5Read(rid, int(ns[i]), float(off[i]), float(sc[i]), list(rows[i]), self)

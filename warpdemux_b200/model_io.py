"""Model parameters of a WarpDemuX DTW-SVM barcode classifier, as plain arrays.

The reference keeps these inside a pickled ``warpdemux.models.dtw_svm.DTW_SVM``
(fields: ``warpdemux/models/dtw_base.py:13-25``, ``dtw_svm.py:26-30``) wrapping
an ``sklearn.svm.SVC(kernel="precomputed", probability=True)``.  The CUDA path
needs only the numbers, laid out the way libsvm walks them
(``sklearn/svm/src/libsvm/svm.cpp:2868-2896``), so this module converts a
reference model object (or its ``.joblib`` file) into :class:`ModelParams` and
stores/loads that as a neutral ``.npz``.

Nothing here computes on the hot path.
"""

from __future__ import annotations

import os
import types
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence

import numpy as np

__all__ = [
    "ModelParams",
    "from_reference_model",
    "load_reference_joblib",
    "load_npz",
    "save_npz",
    "synthetic_model",
    "load_model",
]


@dataclass
class ModelParams:
    """Everything ``DTW_SVM.predict`` reads from the reference model object.

    sv            float64 [n_sv, L]      ``DTW_SVM._X`` (support vectors, class-sorted)
    n_sv_class    int32   [k]            ``SVC._n_support``
    dual_coef     float64 [k-1, n_sv]    ``SVC._dual_coef_``
    rho           float64 [k(k-1)/2]     ``-SVC._intercept_``
    probA, probB  float64 [k(k-1)/2]     ``SVC._probA / _probB``
    thresholds    float64 [k]            ``DTW_SVM.thresholds``
    label_map     int64   [k]            ``DTW_SVM.label_mapper`` (index -> barcode, noise = -1)
    """

    sv: np.ndarray
    n_sv_class: np.ndarray
    dual_coef: np.ndarray
    rho: np.ndarray
    probA: np.ndarray
    probB: np.ndarray
    thresholds: np.ndarray
    label_map: np.ndarray
    window: int = 15
    penalty: float = 0.1
    gamma: float = 1.0
    pwr_dist: int = 1
    block_size: int = 500
    noise_class: bool = True
    name: str = ""

    def __post_init__(self):
        self.sv = np.ascontiguousarray(self.sv, dtype=np.float64)
        self.n_sv_class = np.ascontiguousarray(self.n_sv_class, dtype=np.int32)
        self.dual_coef = np.ascontiguousarray(self.dual_coef, dtype=np.float64)
        self.rho = np.ascontiguousarray(self.rho, dtype=np.float64)
        self.probA = np.ascontiguousarray(self.probA, dtype=np.float64)
        self.probB = np.ascontiguousarray(self.probB, dtype=np.float64)
        self.thresholds = np.ascontiguousarray(self.thresholds, dtype=np.float64)
        self.label_map = np.ascontiguousarray(self.label_map, dtype=np.int64)
        self.validate()

    # -- shape helpers -----------------------------------------------------
    @property
    def n_sv(self) -> int:
        return int(self.sv.shape[0])

    @property
    def L(self) -> int:
        return int(self.sv.shape[1])

    @property
    def k(self) -> int:
        return int(self.n_sv_class.shape[0])

    @property
    def n_pairs(self) -> int:
        return self.k * (self.k - 1) // 2

    @property
    def label_mapper(self) -> Dict[int, int]:
        return {i: int(v) for i, v in enumerate(self.label_map)}

    def band_cells(self) -> int:
        """DTW cells inside the Sakoe-Chiba band for one (read, SV) pair."""
        L, w = self.L, self.window if self.window > 0 else self.L
        return sum(min(L, i + w) - max(0, i - w + 1) for i in range(L))

    def validate(self) -> None:
        k, n_sv = self.k, self.n_sv
        if self.sv.ndim != 2:
            raise ValueError("sv must be [n_sv, L]")
        if k < 2:
            raise ValueError("need at least two classes")
        if int(self.n_sv_class.sum()) != n_sv:
            raise ValueError("n_sv_class does not sum to n_sv")
        if self.dual_coef.shape != (k - 1, n_sv):
            raise ValueError(f"dual_coef must be [{k - 1}, {n_sv}], got {self.dual_coef.shape}")
        for nm in ("rho", "probA", "probB"):
            if getattr(self, nm).shape != (self.n_pairs,):
                raise ValueError(f"{nm} must have length k(k-1)/2 = {self.n_pairs}")
        if self.thresholds.shape != (k,) or self.label_map.shape != (k,):
            raise ValueError("thresholds and label_map must have length k")


def from_reference_model(obj, name: str = "") -> ModelParams:
    """Convert an (unpickled) reference ``DTW_SVM`` object; duck-typed, so it
    works on the real class and on the stub the safe loader substitutes."""
    svc = obj.model
    if svc is None or obj._X is None:
        raise ValueError("Model not trained yet.")  # dtw_svm.py:65-68
    kernel = getattr(svc, "kernel", "precomputed")
    if kernel != "precomputed":
        raise ValueError(f"SVC kernel must be 'precomputed', got {kernel!r}")
    support = np.asarray(svc.support_)
    X = np.asarray(obj._X, dtype=np.float64)
    if support.shape[0] != X.shape[0] or not np.array_equal(support, np.arange(X.shape[0])):
        # libsvm's precomputed kernel reads K[:, support_[s]] (svm.cpp:518-522);
        # gather so that column s of our kernel is SV s.
        X = X[support]
    probA = np.asarray(vars(svc).get("_probA", vars(svc).get("probA_")), dtype=np.float64)
    probB = np.asarray(vars(svc).get("_probB", vars(svc).get("probB_")), dtype=np.float64)
    if probA.size == 0:
        raise ValueError("SVC was not fitted with probability=True")
    dual = vars(svc).get("_dual_coef_", vars(svc).get("dual_coef_"))
    if hasattr(dual, "toarray"):
        dual = dual.toarray()
    intercept = np.asarray(vars(svc).get("_intercept_", vars(svc).get("intercept_")), dtype=np.float64)
    k = int(np.asarray(svc._n_support).shape[0])
    lm = obj.label_mapper
    label_map = np.array([lm[i] for i in range(k)], dtype=np.int64)
    thr = obj.thresholds
    thresholds = np.zeros(k) if thr is None else np.asarray(thr, dtype=np.float64)
    return ModelParams(
        sv=X,
        n_sv_class=np.asarray(svc._n_support, dtype=np.int32),
        dual_coef=dual,
        rho=-intercept,
        probA=probA,
        probB=probB,
        thresholds=thresholds,
        label_map=label_map,
        window=int(obj.window),
        penalty=float(obj.penalty),
        gamma=float(obj.gamma),
        pwr_dist=int(obj.pwr_dist),
        block_size=int(getattr(obj, "block_size", 500) or 500),
        noise_class=bool(getattr(obj, "noise_class", True)),
        name=name,
    )


class _RefModelStub:
    """Stands in for ``warpdemux.models.*`` classes while unpickling a model
    file on a machine where the reference package is not importable."""

    def __setstate__(self, state):
        self.__dict__.update(state)


# Exact (module, name) pairs a DTW_SVM model file may reference.  The five shipped DTW_SVM files need only the first four
# (recorded by instrumenting the unpickler); the rest are the data-only globals other numpy / sklearn versions emit for the
# same objects (array reconstruction, scalars, sparse dual coefficients).  Whole packages are NOT admitted: `numpy` and
# `sklearn.utils` contain callables that execute code given as a string (numpy.testing._private.utils.runstring, ...).
_ALLOWED_GLOBALS = frozenset({
    ("joblib.numpy_pickle", "NumpyArrayWrapper"),
    ("numpy", "dtype"),
    ("numpy", "ndarray"),
    ("sklearn.svm._classes", "SVC"),
    ("numpy.core.multiarray", "_reconstruct"),
    ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "scalar"),
    ("numpy", "float64"),
    ("numpy", "int64"),
    ("numpy", "int32"),
    ("numpy", "bool_"),
    ("scipy.sparse._csr", "csr_matrix"),
    ("scipy.sparse._csr", "csr_array"),
    ("scipy.sparse.csr", "csr_matrix"),
    ("collections", "OrderedDict"),
    ("builtins", "set"), ("builtins", "frozenset"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"),
    ("builtins", "slice"), ("builtins", "complex"), ("builtins", "bytearray"), ("builtins", "object"),
})


def load_reference_joblib(path: str) -> ModelParams:
    """Read a reference ``*.joblib`` model file with an allow-listing unpickler: only the exact globals of
    ``_ALLOWED_GLOBALS`` (array / dtype / SVC constructors, no package-wide admission) are resolved and
    ``warpdemux.models.*`` classes are replaced by a state-holding stub; anything else raises.  That keeps a crafted
    file from reaching an arbitrary callable through a REDUCE opcode; it is still the reference's trust model
    (``joblib.load`` of files shipped with the package) — do not feed it files from untrusted sources."""
    import warnings

    from joblib import numpy_pickle as npk

    class _SafeUnpickler(npk.NumpyUnpickler):
        def find_class(self, module, name):
            if module.startswith("warpdemux.models"):
                return _RefModelStub
            if (module, name) in _ALLOWED_GLOBALS:
                return super().find_class(module, name)
            raise ValueError(f"refusing to unpickle global {module}.{name} from {path}")

    with open(path, "rb") as f, warnings.catch_warnings():
        warnings.simplefilter("ignore")  # sklearn InconsistentVersionWarning
        with npk._validate_fileobject_and_memmap(f, path, None) as (fobj, _):
            obj = _SafeUnpickler(path, fobj, True).load()
    name = os.path.splitext(os.path.basename(path))[0]
    return from_reference_model(obj, name=name)


_NPZ_ARRAYS = ("sv", "n_sv_class", "dual_coef", "rho", "probA", "probB", "thresholds", "label_map")
_NPZ_SCALARS = ("window", "penalty", "gamma", "pwr_dist", "block_size", "noise_class", "name")


def save_npz(params: ModelParams, path: str) -> None:
    d = {k: getattr(params, k) for k in _NPZ_ARRAYS}
    d.update({k: np.asarray(getattr(params, k)) for k in _NPZ_SCALARS})
    np.savez_compressed(path, **d)


def load_npz(path: str) -> ModelParams:
    with np.load(path, allow_pickle=False) as z:
        kw = {k: z[k] for k in _NPZ_ARRAYS}
        kw.update(
            window=int(z["window"]),
            penalty=float(z["penalty"]),
            gamma=float(z["gamma"]),
            pwr_dist=int(z["pwr_dist"]),
            block_size=int(z["block_size"]),
            noise_class=bool(z["noise_class"]),
            name=str(z["name"]),
        )
    return ModelParams(**kw)


def load_model(path: str) -> ModelParams:
    """``.npz`` (this package's neutral format) or a reference ``.joblib``."""
    if path.endswith(".npz"):
        return load_npz(path)
    return load_reference_joblib(path)


def synthetic_model(
    n_sv_class: Sequence[int],
    L: int = 25,
    window: int = 15,
    penalty: float = 0.1,
    gamma: float = 1.0,
    seed: int = 0,
    name: str = "synthetic",
) -> ModelParams:
    """A random model of a given shape (for shape/throughput tests where no
    trained model is at hand).  Class c's support vectors scatter around a
    random class template, are mean/std-normalised like real fingerprints
    (``sig_proc.py:546-552``); coefficients have libsvm's sign structure."""
    rng = np.random.default_rng(seed)
    n_sv_class = np.asarray(n_sv_class, dtype=np.int32)
    k = int(n_sv_class.shape[0])
    n_sv = int(n_sv_class.sum())
    cls = np.repeat(np.arange(k), n_sv_class)
    templates = rng.normal(size=(k, L))
    sv = templates[cls] + 0.6 * rng.normal(size=(n_sv, L))
    sv = (sv - sv.mean(axis=1, keepdims=True)) / sv.std(axis=1, keepdims=True)
    dual = np.zeros((k - 1, n_sv))
    for s in range(n_sv):
        c = cls[s]
        for r in range(k - 1):
            o = r if r < c else r + 1
            mag = rng.uniform(0.0, 2.0) if rng.random() < 0.6 else 0.0
            dual[r, s] = mag if c < o else -mag  # y=+1 for the lower class of the pair
    n_pairs = k * (k - 1) // 2
    return ModelParams(
        sv=sv,
        n_sv_class=n_sv_class,
        dual_coef=dual,
        rho=rng.normal(scale=0.2, size=n_pairs),
        probA=-rng.uniform(1.5, 4.0, size=n_pairs),
        probB=rng.normal(scale=0.2, size=n_pairs),
        thresholds=np.concatenate([rng.uniform(0.1, 0.8, size=k - 1), [0.0]]),
        label_map=np.concatenate([np.arange(1, k), [-1]]).astype(np.int64),
        window=window,
        penalty=penalty,
        gamma=gamma,
        name=name,
    )

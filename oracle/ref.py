"""ctypes view of oracle/_ref/libmscnn_ref.so (the verbatim-compiled reference CPU layers).

TEST INFRASTRUCTURE: see oracle/__init__.py.  `RefNet` mirrors the slice of caffe::Net the
MATLAB/pycaffe drivers use (net.cpp:27-214): construct from a deploy prototxt, copy params by
layer name, set input data, forward, read blobs.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_LIB_PATH = Path(__file__).resolve().parent / "_ref" / "libmscnn_ref.so"
_lib = None


def available() -> bool:
    return _LIB_PATH.exists()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{_LIB_PATH} missing: run `python oracle/build_ref.py` where "
                               "/root/reference is mounted")
        L = C.CDLL(str(_LIB_PATH))
        L.mscnn_ref_net_create.restype = C.c_void_p
        L.mscnn_ref_net_create.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.mscnn_ref_net_destroy.argtypes = [C.c_void_p]
        L.mscnn_ref_net_num_layers.argtypes = [C.c_void_p]
        for f in ("mscnn_ref_net_layer_name", "mscnn_ref_net_layer_type"):
            getattr(L, f).restype = C.c_char_p
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        L.mscnn_ref_net_num_params.argtypes = [C.c_void_p, C.c_char_p]
        L.mscnn_ref_net_param_shape.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        L.mscnn_ref_net_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_long]
        L.mscnn_ref_net_blob_shape.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.mscnn_ref_net_reshape_blob.argtypes = [C.c_void_p, C.c_char_p] + [C.c_int] * 4
        L.mscnn_ref_net_set_blob.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long]
        L.mscnn_ref_net_get_blob.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long]
        L.mscnn_ref_net_forward.restype = C.c_double
        L.mscnn_ref_net_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.mscnn_ref_blas_backend.restype = C.c_char_p
        L.mscnn_ref_blas_threads.restype = C.c_int
        L.mscnn_ref_blas_set_threads.restype = C.c_int
        L.mscnn_ref_blas_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def blas_backend() -> str:
    return lib().mscnn_ref_blas_backend().decode()


def blas_threads() -> int:
    return int(lib().mscnn_ref_blas_threads())


def set_blas_threads(n: int) -> int:
    """Pin the BLAS thread count explicitly (torchrun exports OMP_NUM_THREADS=1); returns the count in effect."""
    return int(lib().mscnn_ref_blas_set_threads(int(n)))


class RefNet:
    def __init__(self, prototxt: str, is_path: bool | None = None, batch: int = 0):
        if is_path is None:
            is_path = "\n" not in prototxt and Path(prototxt).exists()
        self._h = lib().mscnn_ref_net_create(str(prototxt).encode(), int(is_path), int(batch))
        if not self._h:
            raise RuntimeError("reference net construction failed")
        n = lib().mscnn_ref_net_num_layers(self._h)
        self.layer_names = [lib().mscnn_ref_net_layer_name(self._h, i).decode() for i in range(n)]
        self.layer_types = [lib().mscnn_ref_net_layer_type(self._h, i).decode() for i in range(n)]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().mscnn_ref_net_destroy(self._h)
            self._h = None

    # -- params ------------------------------------------------------------------------
    def param_shapes(self, layer: str) -> list[tuple[int, ...]]:
        n = lib().mscnn_ref_net_num_params(self._h, layer.encode())
        out = []
        for i in range(max(n, 0)):
            s = (C.c_int * 4)()
            nd = lib().mscnn_ref_net_param_shape(self._h, layer.encode(), i, s)
            out.append(tuple(s[:nd]))
        return out

    def set_param(self, layer: str, idx: int, arr: np.ndarray) -> None:
        a = np.ascontiguousarray(arr, dtype=np.float32)
        rc = lib().mscnn_ref_net_set_param(self._h, layer.encode(), idx, a.ctypes.data, a.size)
        if rc:
            raise ValueError(f"set_param({layer},{idx}) rc={rc} (size {a.size} vs {self.param_shapes(layer)})")

    def set_params(self, weights: dict[str, list[np.ndarray]]) -> None:
        """Copy by layer name, like Net::CopyTrainedLayersFrom (net.cpp:750-785)."""
        for name, blobs in weights.items():
            if name in self.layer_names:
                for i, b in enumerate(blobs):
                    self.set_param(name, i, b)

    # -- blobs -------------------------------------------------------------------------
    def blob_shape(self, name: str) -> tuple[int, ...]:
        s = (C.c_int * 4)()
        nd = lib().mscnn_ref_net_blob_shape(self._h, name.encode(), s)
        if nd < 0:
            raise KeyError(name)
        return tuple(s[:nd])

    def set_blob(self, name: str, arr: np.ndarray) -> None:
        a = np.ascontiguousarray(arr, dtype=np.float32)
        shp = tuple(a.shape) + (1,) * (4 - a.ndim)
        if lib().mscnn_ref_net_reshape_blob(self._h, name.encode(), *shp):
            raise KeyError(name)
        rc = lib().mscnn_ref_net_set_blob(self._h, name.encode(), a.ctypes.data, a.size)
        if rc:
            raise ValueError(f"set_blob({name}) rc={rc}")

    def blob(self, name: str) -> np.ndarray:
        shp = self.blob_shape(name)
        out = np.empty(shp, dtype=np.float32)
        rc = lib().mscnn_ref_net_get_blob(self._h, name.encode(), out.ctypes.data, out.size)
        if rc:
            raise ValueError(f"get_blob({name}) rc={rc}")
        return out

    # -- forward -----------------------------------------------------------------------
    def forward(self, start: str | None = None, end: str | None = None) -> dict[str, float]:
        """Run layers [start, end] (names, inclusive); returns per-layer wall-clock ms."""
        i0 = self.layer_names.index(start) if start else 0
        i1 = self.layer_names.index(end) if end else len(self.layer_names) - 1
        ms = (C.c_double * len(self.layer_names))()
        lib().mscnn_ref_net_forward(self._h, i0, i1, ms)
        return {self.layer_names[i]: ms[i] for i in range(i0, i1 + 1)}

"""ctypes loader of the plain-C integer oracle ``oracle/restate_int.c`` (TEST INFRASTRUCTURE ONLY).

Built by ``make -C oracle`` (``__graft_entry__.build()`` does it) into ``oracle/_build/liboracle_int.so``.
The numpy forms in ``oracle/restate.py`` and these C functions were written independently from the
same reference lines; the CPU tests check them against each other and against the fixtures."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle_int.so")
_lib = None


def load(build: bool = True) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(_HERE, "restate_int.c")
    if build and (not os.path.isfile(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    lib = C.CDLL(LIB_PATH)
    lib.oracle_grid_num_edges.restype = C.c_int64
    lib.oracle_grid_num_edges.argtypes = [C.c_int, C.c_int]
    lib.oracle_grid_edge_index.argtypes = [C.c_int, C.c_int, C.c_void_p]
    lib.oracle_complete_edge_index.argtypes = [C.c_int, C.c_void_p]
    lib.oracle_csr_from_coo.restype = C.c_int64
    lib.oracle_csr_from_coo.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_argmax_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.oracle_nearest_index.argtypes = [C.c_int, C.c_int, C.c_void_p]
    lib.oracle_unpool_nearest.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    _lib = lib
    return lib


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def grid_edge_index(Hp: int, Wp: int) -> np.ndarray:
    lib = load()
    E = int(lib.oracle_grid_num_edges(Hp, Wp))
    ei = np.zeros((2, E), dtype=np.int64)
    if E:
        lib.oracle_grid_edge_index(Hp, Wp, _p(ei))
    return ei


def complete_edge_index(K: int) -> np.ndarray:
    ei = np.zeros((2, max(K * (K - 1), 0)), dtype=np.int64)
    if ei.size:
        load().oracle_complete_edge_index(K, _p(ei))
    return ei


def csr_from_coo(ei: np.ndarray, N: int, by_target: bool = True):
    ei = np.ascontiguousarray(ei, dtype=np.int64)
    E = ei.shape[1]
    rowptr = np.zeros(N + 1, dtype=np.int32)
    col = np.zeros(E, dtype=np.int32)
    eid = np.zeros(E, dtype=np.int32)
    bad = int(load().oracle_csr_from_coo(_p(ei), E, N, int(by_target), _p(rowptr), _p(col), _p(eid)))
    return rowptr, col, eid, bad


def argmax_rows(S: np.ndarray) -> np.ndarray:
    S = np.ascontiguousarray(S, dtype=np.float32)
    out = np.zeros(S.shape[0], dtype=np.int32)
    load().oracle_argmax_rows(_p(S), S.shape[0], S.shape[1], _p(out))
    return out


def nearest_index(out_size: int, in_size: int) -> np.ndarray:
    idx = np.zeros(out_size, dtype=np.int32)
    load().oracle_nearest_index(out_size, in_size, _p(idx))
    return idx


def unpool_nearest(table: np.ndarray, labels, Hp: int, Wp: int, H: int, W: int) -> np.ndarray:
    table = np.ascontiguousarray(table, dtype=np.float32)
    D = table.shape[1]
    out = np.zeros((D, H, W), dtype=np.float32)
    lab = None if labels is None else np.ascontiguousarray(labels, dtype=np.int32)
    load().oracle_unpool_nearest(_p(table), None if lab is None else _p(lab), D, Hp, Wp, H, W, _p(out))
    return out

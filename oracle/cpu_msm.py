"""ctypes wrapper of oracle/libcpu_msm.so (cpu_msm.c) -- TEST INFRASTRUCTURE ONLY.
Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libcpu_msm.so")
_lib = None
NO_INF = C.c_size_t(-1).value


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-C", _HERE], check=True)
        _lib = C.CDLL(_PATH)
        _lib.oracle_msm.restype = C.c_int
        _lib.oracle_msm.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                    C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        _lib.oracle_ark_window.restype = C.c_int
        _lib.oracle_ark_window.argtypes = [C.c_size_t]
        _lib.oracle_dlog_checksum.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_scalar_mul_gen.argtypes = [C.c_void_p, C.c_void_p]
        _lib.oracle_jac_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_testkit_generate.restype = C.c_int
        _lib.oracle_testkit_generate.argtypes = [C.c_uint64, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def msm(bases: np.ndarray, scalars: np.ndarray, threads: int = 0, window_bits: int = 0):
    """bases (n, 8|9) u64 arkworks records, scalars (n, 4) u64 Montgomery -> (12 u64 Jacobian, threads used)."""
    assert bases.dtype == np.uint64 and scalars.dtype == np.uint64
    bases = np.ascontiguousarray(bases)
    scalars = np.ascontiguousarray(scalars)
    n = min(len(bases), len(scalars))
    stride = bases.shape[1] * 8
    inf_off = 64 if bases.shape[1] == 9 else NO_INF
    out = np.zeros(12, dtype=np.uint64)
    used = lib().oracle_msm(bases.ctypes.data, stride, 0, 32, inf_off, scalars.ctypes.data, 32, n,
                            threads or (os.cpu_count() or 1), window_bits, out.ctypes.data)
    return out, used


def dlog_checksum(scalars_mont: np.ndarray, t1: np.ndarray, t2: np.ndarray) -> np.ndarray:
    scalars_mont = np.ascontiguousarray(scalars_mont)
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_dlog_checksum(scalars_mont.ctypes.data, len(scalars_mont), np.ascontiguousarray(t1).ctypes.data,
                               np.ascontiguousarray(t2).ctypes.data, out.ctypes.data)
    return out


def scalar_mul_gen(k_words: np.ndarray) -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    lib().oracle_scalar_mul_gen(np.ascontiguousarray(k_words).ctypes.data, out.ctypes.data)
    return out


def jac_add(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    lib().oracle_jac_add(np.ascontiguousarray(a).ctypes.data, np.ascontiguousarray(b).ctypes.data, out.ctypes.data)
    return out


def testkit_generate(seed: int, n: int, want_bases: bool = True):
    """CPU restatement of the device test kit's generator (same bytes): (bases (n, 8) u64 | None, scalars (n, 4) u64,
    t1 dlogs (4096, 4), t2 dlogs (ceil(n/4096), 4))."""
    n2 = (n + 4095) // 4096
    bases = np.zeros((n, 8), dtype=np.uint64) if want_bases else None
    scalars = np.zeros((n, 4), dtype=np.uint64)
    t1 = np.zeros((4096, 4), dtype=np.uint64)
    t2 = np.zeros((n2, 4), dtype=np.uint64)
    rc = lib().oracle_testkit_generate(seed, n, bases.ctypes.data if want_bases else None, scalars.ctypes.data,
                                       t1.ctypes.data, t2.ctypes.data)
    assert rc == 0
    return bases, scalars, t1, t2

"""ctypes binding of libb200msm.so + the Python mirror of the reference's public entry point.

Reference interface mirrored (same name pattern, argument meaning and error behaviour):

    pub fn metal_variable_base_msm(bases: &[G1Affine], scalars: &[Fr])
        -> Result<G1Projective, Box<dyn Error>>
    /root/reference/mopro-msm/src/msm/metal_msm/metal_msm.rs:642-695

    * empty input            -> Err("Empty input")            (:647-649)  -> raises MsmError("Empty input")
    * bases.len != scalars.len -> truncate to the shorter     (:652-656)  -> same
    * result == G::msm(bases, scalars) as a group element     (tests/cuzk/e2e.rs:58-61)

Memory model: `bases` is an (n, 9) uint64 array = n arkworks `G1Affine` records of 72 bytes
(x: 4 u64 Montgomery LE, y: 4 u64, infinity: bool in the low byte of the 9th word), or an
(n, 8) array (64-byte records, no infinity flag); `scalars` is an (n, 4) uint64 array of `Fr`
Montgomery words.  There is NO CPU path in this module: if the CUDA library is missing or no
B200 is visible, loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200MSM_LIB") or os.path.join(_HERE, "lib", "libb200msm.so")

NO_INF = C.c_size_t(-1).value
_P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


class MsmError(RuntimeError):
    """Maps the C ABI's negative return codes (+ last_error text) like the shim's Box<dyn Error>."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("decompose_ms", C.c_float), ("sort_ms", C.c_float),
                ("accumulate_ms", C.c_float), ("reduce_ms", C.c_float), ("total_ms", C.c_float),
                ("window_bits", C.c_int), ("num_windows", C.c_int),
                ("entries", C.c_ulonglong), ("kernel_launches", C.c_ulonglong)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None

# every symbol include/b200msm.h declares: (name, restype, argtypes)
_vp, _sz, _i, _u64p = C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint64)
SYMBOLS = {
    "b200msm_create": (_i, [C.POINTER(_vp), C.POINTER(_i), _i]),
    "b200msm_destroy": (None, [_vp]),
    "b200msm_last_error": (C.c_char_p, [_vp]),
    "b200msm_build_id": (C.c_char_p, []),
    "b200msm_device_count": (_i, [_vp]),
    "b200msm_set_option": (_i, [_vp, C.c_char_p, C.c_longlong]),
    "b200msm_last_timings": (_i, [_vp, C.POINTER(Timings)]),
    "b200msm_last_sort_engine": (_i, [_vp]),
    "b200msm_auto_window_bits": (_i, [_vp, _sz]),
    "b200msm_bn254_g1_msm": (_i, [_vp, _vp, _sz, _sz, _sz, _sz, _vp, _sz, _sz, _u64p]),
    "b200msm_bn254_g2_msm": (_i, [_vp, _vp, _sz, _sz, _sz, _sz, _vp, _sz, _sz, _u64p]),
    "b200msm_g2_register_bases": (_i, [_vp, _vp, _sz, _sz, _sz, _sz, _sz, _i, C.POINTER(_vp)]),
    "b200msm_g2_release_bases": (_i, [_vp, _vp]),
    "b200msm_g2_msm_registered": (_i, [_vp, _vp, _vp, _sz, _sz, _u64p]),
    "b200msm_register_bases": (_i, [_vp, _vp, _sz, _sz, _sz, _sz, _sz, C.POINTER(_vp)]),
    "b200msm_register_bases_on": (_i, [_vp, _vp, _sz, _sz, _sz, _sz, _sz, C.POINTER(_i), _i, C.POINTER(_vp)]),
    "b200msm_register_bases_ex": (_i, [_vp, _vp, _sz, _sz, _sz, _sz, _sz, C.POINTER(_i), _i, _i, C.POINTER(_vp)]),
    "b200msm_release_bases": (_i, [_vp, _vp]),
    "b200msm_bases_len": (_sz, [_vp]),
    "b200msm_msm_registered": (_i, [_vp, _vp, _vp, _sz, _sz, _u64p]),
    "b200msm_msm_batch": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_sz), _vp]),
    "b200msm_msm_device": (_i, [_vp, _i, _vp, _vp, _vp, _sz, _vp, _i]),
    "b200msm_sum_partials_device": (_i, [_vp, _i, _vp, _i, _vp, _i]),
    "b200msm_sync": (_i, [_vp]),
    "b200msm_stream": (_vp, [_vp, _i]),
    "b200msm_set_stream": (_i, [_vp, _i, _vp]),
    "b200msm_testkit_occupy_sms": (_i, [_vp, _i, _i, C.c_double]),
    "b200msm_testkit_release_sms": (_i, [_vp, _i]),
    "b200msm_testkit_imad_peak": (_i, [_vp, _i, C.POINTER(C.c_double)]),
    "b200msm_decompress_g1": (_i, [_vp, _vp, _sz, _vp, C.POINTER(C.c_uint64)]),
    "b200msm_fr_to_montgomery": (_i, [_vp, _vp, _sz, _vp]),
    "b200msm_testkit_generate": (_i, [_vp, _i, C.c_uint64, _sz, _vp, _vp, _vp, _vp]),
    "b200msm_testkit_op": (_i, [_vp, _i, _vp, _vp, _vp, _sz]),
    "b200math_apply": (_i, [_vp, _i, _vp, _vp, _vp, _sz]),
    "b200msm_testkit_window_sums": (_i, [_vp, _vp, _vp, _sz, _i, _vp, C.POINTER(_i)]),
    "b200msm_testkit_g2_window_sums": (_i, [_vp, _vp, _vp, _sz, _i, _vp, C.POINTER(_i)]),
    "b200msm_testkit_slice_plan": (_i, [_sz, _i, _i, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_i)]),
    "b200msm_testkit_parallel_copy": (_i, [_vp, _vp, _sz, _i]),
    "b200msm_testkit_table": (_i, [_vp, _vp, _i, _sz, _vp, C.POINTER(_i), C.POINTER(_i)]),
    "b200msm_testkit_sort": (_i, [_vp, _vp, _sz, _i, _vp, _vp, C.POINTER(C.c_uint64), C.POINTER(_i), C.POINTER(C.c_uint64)]),
}


def _source_build_id() -> Optional[str]:
    """Hash of csrc/ + include/ in this checkout (None when the sources are not beside the module)."""
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("b200msm_build_id", os.path.join(_HERE, "build_id.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.build_id() if mod.source_files() else None
    except Exception:
        return None


def _bind(p: str):
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def load_library(path: Optional[str] = None):
    """Load libb200msm.so and bind every declared symbol.  Raises if the library is missing:
    there is deliberately no fallback.  The library carries a hash of the sources it was compiled
    from; if it differs from the checkout's (`build_id.py`), the library is rebuilt once with
    `make` and, if it still differs, loading FAILS: a stale binary is never what gets tested."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p) and path is None and os.path.exists(os.path.join(_HERE, "Makefile")):
        import subprocess
        subprocess.run(["make", "-C", _HERE, "lib/libb200msm.so"], check=False, capture_output=True)
    if not os.path.exists(p):
        raise MsmError(-4, f"{p} not found: build it with `make -C gpu-acceleration_b200` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    lib = _bind(p)
    want = _source_build_id() if path is None and not os.environ.get("B200MSM_LIB") else None
    if want is not None:
        have = lib.b200msm_build_id().decode()
        if have != want:
            # dlopen caches by path: build to a fresh name so that the new code is what gets mapped
            import subprocess
            r = subprocess.run(["make", "-C", _HERE, "lib/libb200msm.so"], capture_output=True, text=True)
            fresh = os.path.join(_HERE, "lib", f"libb200msm.{want}.so")
            if r.returncode == 0:
                import shutil
                shutil.copyfile(p, fresh)
                lib = _bind(fresh)
                have = lib.b200msm_build_id().decode()
            if have != want:
                raise MsmError(-4, f"{p} was built from other sources (library {have}, checkout {want}) and the rebuild "
                                   f"failed: run `make -C gpu-acceleration_b200`\n{r.stderr[-2000:]}")
    if path is None:
        _lib = lib
    return lib


def _ptr(a) -> int:
    """Host numpy array / bytes-like / torch tensor / int -> raw address."""
    if a is None:
        return 0
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


@dataclass
class G1Projective:
    """arkworks `G1Projective {x, y, z}`: Jacobian, Montgomery words.  `==` is arkworks'
    cross-multiplied projective equality (what the reference's assert_eq! uses)."""
    words: np.ndarray  # (12,) uint64

    def _ints(self):
        rinv = pow(1 << 256, -1, _P)
        vals = []
        for c in range(3):
            v = 0
            for j in range(4):
                v |= int(self.words[4 * c + j]) << (64 * j)
            vals.append(v * rinv % _P)
        return vals

    def is_zero(self) -> bool:
        return all(int(w) == 0 for w in self.words[8:12])

    def into_affine(self):
        """(x, y) canonical ints, or None for the identity."""
        x, y, z = self._ints()
        if z == 0:
            return None
        zi = pow(z, -1, _P)
        return (x * zi * zi % _P, y * zi * zi * zi % _P)

    def __eq__(self, other):
        if not isinstance(other, G1Projective):
            return NotImplemented
        return self.into_affine() == other.into_affine()


class Bases:
    def __init__(self, ctx: "Context", handle: int):
        self.ctx, self.handle = ctx, handle

    def __len__(self):
        return self.ctx.lib.b200msm_bases_len(self.handle)

    def release(self):
        if self.handle:
            self.ctx._check(self.ctx.lib.b200msm_release_bases(self.ctx.h, self.handle))
            self.handle = 0


def _base_layout(bases: np.ndarray):
    assert bases.dtype == np.uint64 and bases.ndim == 2 and bases.shape[1] in (8, 9), \
        "bases must be (n, 9) [x, y, infinity] or (n, 8) [x, y] uint64"
    stride = bases.shape[1] * 8
    return stride, 0, 32, (64 if bases.shape[1] == 9 else NO_INF)


class Context:
    """Persistent engine context (replaces the per-call MetalMSMPipeline, metal_msm.rs:48-62)."""

    def __init__(self, devices: Optional[Sequence[int]] = None):
        self.lib = load_library()
        self.h = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.b200msm_create(C.byref(self.h), arr, len(devices))
        else:
            rc = self.lib.b200msm_create(C.byref(self.h), None, 0)
        if rc != 0:
            raise MsmError(rc, self.lib.b200msm_last_error(None).decode())

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.b200msm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise MsmError(rc, self.lib.b200msm_last_error(self.h).decode())

    @property
    def device_count(self) -> int:
        return self.lib.b200msm_device_count(self.h)

    def set_option(self, key: str, value: int):
        self._check(self.lib.b200msm_set_option(self.h, key.encode(), value))

    def timings(self) -> dict:
        t = Timings()
        self._check(self.lib.b200msm_last_timings(self.h, C.byref(t)))
        return t.as_dict()

    def auto_window_bits(self, n: int) -> int:
        return self.lib.b200msm_auto_window_bits(self.h, n)

    def last_sort_engine(self) -> int:
        """K1 + K2 engine of the most recent MSM: 0 cursor atomics, 1 ranked, 2 partitioned."""
        return self.lib.b200msm_last_sort_engine(self.h)

    # ---- host-buffer drop-in call
    def msm(self, bases: np.ndarray, scalars: np.ndarray, n: Optional[int] = None) -> G1Projective:
        stride, xo, yo, io = _base_layout(bases)
        assert scalars.dtype == np.uint64 and scalars.ndim == 2 and scalars.shape[1] == 4
        n = min(len(bases), len(scalars)) if n is None else n
        out = np.zeros(12, dtype=np.uint64)
        self._check(self.lib.b200msm_bn254_g1_msm(self.h, _ptr(bases), stride, xo, yo, io, _ptr(scalars), 32, n,
                                                  out.ctypes.data_as(_u64p)))
        return G1Projective(out)

    def msm_raw(self, bases_ptr: int, stride: int, x_off: int, y_off: int, inf_off: int, scalars_ptr: int,
                scalar_stride: int, n: int) -> G1Projective:
        out = np.zeros(12, dtype=np.uint64)
        self._check(self.lib.b200msm_bn254_g1_msm(self.h, bases_ptr, stride, x_off, y_off, inf_off, scalars_ptr,
                                                  scalar_stride, n, out.ctypes.data_as(_u64p)))
        return G1Projective(out)

    # ---- registered bases
    def msm_g2(self, bases: np.ndarray, scalars: np.ndarray) -> np.ndarray:
        """BN254 G2 MSM.  bases: (n, 17) uint64 [x.c0, x.c1, y.c0, y.c1, infinity] or (n, 16) without the flag word;
        returns the 24 Jacobian words (G2Projective memory)."""
        assert bases.dtype == np.uint64 and bases.ndim == 2 and bases.shape[1] in (16, 17)
        n = min(len(bases), len(scalars))
        out = np.zeros(24, dtype=np.uint64)
        self._check(self.lib.b200msm_bn254_g2_msm(self.h, _ptr(bases), bases.shape[1] * 8, 0, 64,
                                                  128 if bases.shape[1] == 17 else NO_INF, _ptr(scalars), 32, n,
                                                  out.ctypes.data_as(_u64p)))
        return out

    def g2_register_bases(self, bases: np.ndarray, precompute: int = 0) -> int:
        """-> opaque handle (int) for g2_msm_registered / g2_release_bases."""
        assert bases.dtype == np.uint64 and bases.ndim == 2 and bases.shape[1] in (16, 17)
        h = C.c_void_p()
        self._check(self.lib.b200msm_g2_register_bases(self.h, _ptr(bases), bases.shape[1] * 8, 0, 64,
                                                       128 if bases.shape[1] == 17 else NO_INF, len(bases), precompute, C.byref(h)))
        return h.value

    def g2_msm_registered(self, handle: int, scalars: np.ndarray) -> np.ndarray:
        out = np.zeros(24, dtype=np.uint64)
        self._check(self.lib.b200msm_g2_msm_registered(self.h, handle, _ptr(scalars), 32, len(scalars), out.ctypes.data_as(_u64p)))
        return out

    def g2_release_bases(self, handle: int) -> None:
        self._check(self.lib.b200msm_g2_release_bases(self.h, handle))

    def register_bases(self, bases: np.ndarray, dev_indices: Optional[Sequence[int]] = None,
                       precompute: Optional[int] = None) -> Bases:
        """precompute: None = the context's "precompute" option; 0 / 1 / 8..24 = b200msm_register_bases_ex."""
        stride, xo, yo, io = _base_layout(bases)
        h = C.c_void_p()
        if precompute is not None:
            arr = (C.c_int * len(dev_indices))(*dev_indices) if dev_indices else None
            self._check(self.lib.b200msm_register_bases_ex(self.h, _ptr(bases), stride, xo, yo, io, len(bases), arr,
                                                           len(dev_indices) if dev_indices else 0, precompute, C.byref(h)))
        elif dev_indices:
            arr = (C.c_int * len(dev_indices))(*dev_indices)
            self._check(self.lib.b200msm_register_bases_on(self.h, _ptr(bases), stride, xo, yo, io, len(bases), arr,
                                                           len(dev_indices), C.byref(h)))
        else:
            self._check(self.lib.b200msm_register_bases(self.h, _ptr(bases), stride, xo, yo, io, len(bases), C.byref(h)))
        return Bases(self, h.value)

    def msm_registered(self, bases: Bases, scalars: np.ndarray) -> G1Projective:
        out = np.zeros(12, dtype=np.uint64)
        self._check(self.lib.b200msm_msm_registered(self.h, bases.handle, _ptr(scalars), 32, len(scalars),
                                                    out.ctypes.data_as(_u64p)))
        return G1Projective(out)

    def msm_batch(self, bases: Sequence[Bases], scalars: Sequence[np.ndarray]):
        k = len(bases)
        hs = (C.c_void_p * k)(*[b.handle for b in bases])
        sc = (C.c_void_p * k)(*[_ptr(s) for s in scalars])
        ns = (C.c_size_t * k)(*[len(s) for s in scalars])
        out = np.zeros((k, 12), dtype=np.uint64)
        self._check(self.lib.b200msm_msm_batch(self.h, k, hs, sc, ns, out.ctypes.data))
        return [G1Projective(out[i].copy()) for i in range(k)]

    # ---- device-pointer path (torch tensors or raw addresses)
    def msm_device(self, d_bases, d_scalars, n: int, d_out, d_inf_mask=None, dev_index: int = 0, sync: bool = True):
        self._check(self.lib.b200msm_msm_device(self.h, dev_index, _ptr(d_bases), _ptr(d_inf_mask), _ptr(d_scalars), n,
                                                _ptr(d_out), 1 if sync else 0))

    def sum_partials_device(self, d_partials, count: int, d_out, dev_index: int = 0, sync: bool = True):
        self._check(self.lib.b200msm_sum_partials_device(self.h, dev_index, _ptr(d_partials), count, _ptr(d_out),
                                                         1 if sync else 0))

    def sync(self):
        self._check(self.lib.b200msm_sync(self.h))

    def stream(self, dev_index: int = 0) -> int:
        return self.lib.b200msm_stream(self.h, dev_index) or 0

    def set_stream(self, stream_ptr: int, dev_index: int = 0):
        self._check(self.lib.b200msm_set_stream(self.h, dev_index, stream_ptr))

    def imad_peak(self, dev_index: int = 0) -> float:
        v = C.c_double()
        self._check(self.lib.b200msm_testkit_imad_peak(self.h, dev_index, C.byref(v)))
        return v.value

    # ---- benchmark-instance files (reference: src/msm/utils/preprocess.rs)
    def decompress_g1(self, compressed: np.ndarray):
        """compressed: (n, 32) uint8 -> ((n, 8) uint64 Montgomery x||y, number of invalid records)."""
        assert compressed.dtype == np.uint8 and compressed.ndim == 2 and compressed.shape[1] == 32
        out = np.zeros((len(compressed), 8), dtype=np.uint64)
        bad = C.c_uint64()
        self._check(self.lib.b200msm_decompress_g1(self.h, _ptr(np.ascontiguousarray(compressed)), len(compressed), _ptr(out),
                                                   C.byref(bad)))
        return out, bad.value

    def fr_to_montgomery(self, canonical: np.ndarray) -> np.ndarray:
        """(n, 4) uint64 canonical scalars (`BigInt<4>`) -> (n, 4) uint64 `Fr` Montgomery words."""
        assert canonical.dtype == np.uint64 and canonical.ndim == 2 and canonical.shape[1] == 4
        out = np.zeros_like(canonical)
        self._check(self.lib.b200msm_fr_to_montgomery(self.h, _ptr(np.ascontiguousarray(canonical)), len(canonical), _ptr(out)))
        return out

    # ---- test kit
    def testkit_generate(self, seed: int, n: int, d_bases, d_scalars, want_dlogs: bool = False, dev_index: int = 0):
        t1 = np.zeros((4096, 4), dtype=np.uint64) if want_dlogs else None
        t2 = np.zeros(((n + 4095) // 4096, 4), dtype=np.uint64) if want_dlogs else None
        self._check(self.lib.b200msm_testkit_generate(self.h, dev_index, seed, n, _ptr(d_bases), _ptr(d_scalars),
                                                      _ptr(t1), _ptr(t2)))
        return t1, t2

    def occupy_sms(self, n_sms: int, max_seconds: float = 30.0, dev_index: int = 0):
        self._check(self.lib.b200msm_testkit_occupy_sms(self.h, dev_index, n_sms, max_seconds))

    def release_sms(self, dev_index: int = 0):
        self._check(self.lib.b200msm_testkit_release_sms(self.h, dev_index))

    def testkit_op(self, op: int, a: np.ndarray, b: Optional[np.ndarray], out_words: int) -> np.ndarray:
        """Element-wise field / curve operation through the public math library (include/b200math.h: b200math_apply;
        `op` is a b200math_op value)."""
        count = a.shape[0]
        out = np.zeros((count, out_words), dtype=np.uint64)
        self._check(self.lib.b200math_apply(self.h, op, _ptr(a), _ptr(b), _ptr(out), count))
        return out

    math_apply = testkit_op

    def testkit_window_sums(self, bases64: np.ndarray, scalars: np.ndarray, window_bits: int) -> np.ndarray:
        out = np.zeros((64, 16), dtype=np.uint64)
        nw = C.c_int()
        self._check(self.lib.b200msm_testkit_window_sums(self.h, _ptr(bases64), _ptr(scalars), len(scalars), window_bits,
                                                         _ptr(out), C.byref(nw)))
        return out[:nw.value]

    def testkit_g2_window_sums(self, bases128: np.ndarray, scalars: np.ndarray, window_bits: int) -> np.ndarray:
        out = np.zeros((64, 32), dtype=np.uint64)
        nw = C.c_int()
        self._check(self.lib.b200msm_testkit_g2_window_sums(self.h, _ptr(bases128), _ptr(scalars), len(scalars), window_bits,
                                                            _ptr(out), C.byref(nw)))
        return out[:nw.value]

    def testkit_table(self, bases: "Bases", window: int, count: int):
        """-> (records[count, 8] of table window `window`, window_bits, num_windows) of a handle registered with "precompute"."""
        out = np.zeros((count, 8), dtype=np.uint64)
        c, nw = C.c_int(), C.c_int()
        self._check(self.lib.b200msm_testkit_table(self.h, bases.handle, window, count, _ptr(out), C.byref(c), C.byref(nw)))
        return out, c.value, nw.value

    def testkit_sort(self, scalars: np.ndarray, window_bits: int):
        """-> (ends[W, nb], entries, n_pseudo): K1+K2 output for the context's current "glv" option."""
        n = len(scalars)
        nb = (1 << (window_bits - 1)) + 1
        ends = np.zeros(64 * nb, dtype=np.uint32)
        cap = max(2 * n * ((127 + window_bits - 1) // window_bits + 1), n * ((254 + window_bits - 1) // window_bits + 1))
        entries = np.zeros(cap, dtype=np.uint32)
        cnt, nw, npseudo = C.c_uint64(), C.c_int(), C.c_uint64()
        self._check(self.lib.b200msm_testkit_sort(self.h, _ptr(scalars), n, window_bits, _ptr(ends), _ptr(entries),
                                                  C.byref(cnt), C.byref(nw), C.byref(npseudo)))
        return ends[:nw.value * nb].reshape(nw.value, nb), entries[:cnt.value], npseudo.value


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    """Process-global lazily created context behind the zero-config call."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def cuda_variable_base_msm(bases: np.ndarray, scalars: np.ndarray, ctx: Optional[Context] = None) -> G1Projective:
    """Drop-in for `metal_variable_base_msm(&bases, &scalars)` (metal_msm.rs:642-695)."""
    if len(bases) == 0 or len(scalars) == 0:
        raise MsmError(-1, "Empty input")  # metal_msm.rs:647-649
    n = min(len(bases), len(scalars))  # metal_msm.rs:652-656
    return (ctx or default_context()).msm(bases, scalars, n)


# ---------------------------------------------------------------------------------------------------------
# Benchmark-instance files of the reference (src/msm/utils/preprocess.rs): `<dir>/points` and `<dir>/scalars`,
# each a sequence of arkworks `Vec<T>::serialize_compressed` records = u64 LE count + count x 32 bytes.
def read_instance_files(directory: str):
    """Yields (compressed_points (n,32) uint8, canonical_scalars (n,4) uint64) per stored instance
    (FileInputIterator, preprocess.rs:101-131)."""
    with open(os.path.join(directory, "points"), "rb") as fp, open(os.path.join(directory, "scalars"), "rb") as fs:
        while True:
            hp, hs = fp.read(8), fs.read(8)
            if len(hp) < 8 or len(hs) < 8:
                return
            n_p, n_s = int.from_bytes(hp, "little"), int.from_bytes(hs, "little")
            pts = np.frombuffer(fp.read(32 * n_p), dtype=np.uint8)
            sc = np.frombuffer(fs.read(32 * n_s), dtype=np.uint64)
            if len(pts) != 32 * n_p or len(sc) != 4 * n_s:
                return
            yield pts.reshape(n_p, 32).copy(), sc.reshape(n_s, 4).copy()


_R_ORDER = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def write_instance_files(directory: str, instances, append: bool = False):
    """The writer side of the reference's instance format (gen_vectors / serialize_input, preprocess.rs:181-225): appends
    every (bases (n, 8|9) uint64 Montgomery G1Affine records, scalars (n, 4) uint64 Montgomery Fr) instance to
    `<dir>/points` and `<dir>/scalars` in arkworks' compressed serialisation (u64 LE count, then per point 32 bytes:
    canonical LE x, bit 7 of byte 31 = y is the larger of (y, p - y), bit 6 = infinity; per scalar the canonical BigInt).
    Host-side integer work only (Montgomery -> canonical with Python integers): meant for generating benchmark sets."""
    os.makedirs(directory, exist_ok=True)
    rinv_p, rinv_r = pow(1 << 256, -1, _P), pow(1 << 256, -1, _R_ORDER)
    mode = "ab" if append else "wb"
    with open(os.path.join(directory, "points"), mode) as fp, open(os.path.join(directory, "scalars"), mode) as fs:
        for bases, scalars in instances:
            n = min(len(bases), len(scalars))
            fp.write(n.to_bytes(8, "little"))
            fs.write(n.to_bytes(8, "little"))
            for i in range(n):
                b = bases[i]
                xm = sum(int(b[j]) << (64 * j) for j in range(4))
                ym = sum(int(b[4 + j]) << (64 * j) for j in range(4))
                inf = (len(b) == 9 and int(b[8]) & 0xFF) or (xm == 0 and ym == 0)
                if inf:
                    rec = bytearray(32)
                    rec[31] |= 0x40
                else:
                    x, y = xm * rinv_p % _P, ym * rinv_p % _P
                    rec = bytearray(x.to_bytes(32, "little"))
                    if y > _P - y:
                        rec[31] |= 0x80
                fp.write(bytes(rec))
                sm = sum(int(scalars[i][j]) << (64 * j) for j in range(4))
                fs.write((sm * rinv_r % _R_ORDER).to_bytes(32, "little"))


def msm_from_instance(ctx: "Context", compressed_points: np.ndarray, canonical_scalars: np.ndarray) -> G1Projective:
    """benchmark_msm's inner step (arkworks_pippenger.rs:19-29) on the GPU: decode, then MSM."""
    bases, bad = ctx.decompress_g1(compressed_points)
    if bad:
        raise MsmError(-1, f"{bad} records are not valid compressed G1 points")
    scalars = ctx.fr_to_montgomery(canonical_scalars)
    return cuda_variable_base_msm(bases, scalars, ctx)

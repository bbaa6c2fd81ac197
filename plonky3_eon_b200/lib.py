"""ctypes binding of libeon_kzg.so (the C ABI in include/eon_kzg.h).

This is the Python equivalent of the `extern "C"` block a Rust `-sys` crate would hold
(INTEGRATION.md shows that Rust form).  There is NO CPU fallback: if the library is missing
or no sm_100 device is usable, everything raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeon_kzg.so")

EON_OK = 0
EON_ERR_BAD_ARG = -1
EON_ERR_SRS_TOO_SHORT = -2
EON_ERR_CUDA = -3
EON_ERR_OOM = -4
EON_ERR_BAD_HANDLE = -5
EON_ERR_TWO_ADICITY = -6
EON_ERR_BAD_POINT = -7
G1_ENC_HALO2 = 0   # halo2curves >= 0.4 GroupEncoding: sign = byte 31 bit 6, identity = byte 31 bit 7
G1_ENC_LEGACY = 1  # halo2curves <= 0.3: sign = byte 31 bit 7, identity = 32 zero bytes

_u64p = C.c_void_p  # all buffers are passed as raw addresses
_SIGS = {
    "eon_version": (C.c_char_p, []),
    "eon_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "eon_ctx_destroy": (None, [C.c_void_p]),
    "eon_last_error": (C.c_char_p, [C.c_void_p]),
    "eon_ctx_sync": (C.c_int, [C.c_void_p]),
    "eon_ctx_launch_count": (C.c_uint64, [C.c_void_p]),
    "eon_dev_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "eon_dev_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "eon_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "eon_host_free": (C.c_int, [C.c_void_p]),
    "eon_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "eon_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "eon_dft_batch_dev": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t]),
    "eon_coset_dft_batch_dev": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t, _u64p]),
    "eon_idft_batch_dev": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t]),
    "eon_coset_idft_batch_dev": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t, _u64p]),
    "eon_coset_lde_batch_dev": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t, C.c_uint, _u64p]),
    "eon_dft_batch": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t]),
    "eon_coset_dft_batch": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t, _u64p]),
    "eon_idft_batch": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t]),
    "eon_coset_idft_batch": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t, _u64p]),
    "eon_coset_lde_batch": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t, C.c_uint, _u64p]),
    "eon_srs_load_affine": (C.c_int, [C.c_void_p, _u64p, C.c_size_t]),
    "eon_srs_load_compressed": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]),
    "eon_g1_compress": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, _u64p, C.c_int]),
    "eon_g1_decompress": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, _u64p, C.c_int, C.POINTER(C.c_size_t)]),
    "eon_srs_generate_unsafe": (C.c_int, [C.c_void_p, _u64p, C.c_size_t]),
    "eon_srs_size": (C.c_size_t, [C.c_void_p]),
    "eon_srs_read": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, _u64p]),
    "eon_srs_set_window_tables": (C.c_int, [C.c_void_p, C.c_uint]),
    "eon_srs_window_bits": (C.c_uint, [C.c_void_p]),
    "eon_msm_set_rounds": (C.c_int, [C.c_void_p, C.c_int]),
    "eon_msm_rounds_used": (C.c_uint, [C.c_void_p]),
    "eon_msm_window_bits_used": (C.c_uint, [C.c_void_p]),
    "eon_msm_set_sort_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "eon_msm_set_slice_schedule": (C.c_int, [C.c_void_p, C.c_int]),
    "eon_msm_set_split": (C.c_int, [C.c_void_p, C.c_int]),
    "eon_msm_srs_dev": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_size_t, C.c_size_t, _u64p]),
    "eon_msm_srs": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_size_t, C.c_size_t, _u64p]),
    "eon_msm_points": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_size_t, _u64p]),
    "eon_msm_srs_range_dev": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, _u64p]),
    "eon_g1_sum": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, _u64p]),
    "eon_kzg_commit": (C.c_int, [C.c_void_p, _u64p, C.c_uint, C.c_size_t, _u64p, _u64p, C.POINTER(C.c_uint64)]),
    "eon_kzg_commit_lde": (C.c_int, [C.c_void_p, _u64p, C.c_uint, C.c_size_t, _u64p, _u64p, C.POINTER(C.c_uint64),
                                     C.c_uint, _u64p, _u64p]),
    "eon_kzg_commit_lde_dev": (C.c_int, [C.c_void_p, _u64p, C.c_uint, C.c_size_t, _u64p, _u64p, C.POINTER(C.c_uint64),
                                         C.c_uint, _u64p, _u64p]),
    "eon_kzg_commit_dev": (C.c_int, [C.c_void_p, _u64p, C.c_uint, C.c_size_t, _u64p, _u64p, C.POINTER(C.c_uint64)]),
    "eon_kzg_commit_coeffs": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_size_t, _u64p, C.POINTER(C.c_uint64)]),
    "eon_kzg_commit_coeffs_dev": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_size_t, _u64p, C.POINTER(C.c_uint64)]),
    # handle / point-count arrays are numpy uint64 buffers (eon_handle and size_t are both 64-bit here)
    "eon_kzg_commit_quotient": (C.c_int, [C.c_void_p, _u64p, C.c_uint, C.c_size_t, C.c_uint, _u64p, _u64p, _u64p]),
    "eon_kzg_commit_quotient_dev": (C.c_int, [C.c_void_p, _u64p, C.c_uint, C.c_size_t, C.c_uint, _u64p, _u64p, _u64p]),
    "eon_kzg_open_batch": (C.c_int, [C.c_void_p, C.c_size_t, _u64p, _u64p, _u64p, _u64p, _u64p]),
    "eon_kzg_read_coeffs": (C.c_int, [C.c_void_p, C.c_uint64, _u64p]),
    "eon_kzg_evals_on_coset": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint, _u64p, _u64p]),
    "eon_kzg_evals_on_coset_dev": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint, _u64p, _u64p]),
    "eon_kzg_open": (C.c_int, [C.c_void_p, C.c_uint64, _u64p, C.c_size_t, _u64p, _u64p]),
    "eon_handle_dims": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint), C.POINTER(C.c_size_t)]),
    "eon_handle_free": (C.c_int, [C.c_void_p, C.c_uint64]),
    "eon_quotient_and_eval_dev": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_size_t, _u64p, _u64p, _u64p]),
    "eon_coset_lde_batch_ld": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, _u64p, C.c_size_t, C.c_uint, C.c_size_t, C.c_uint,
                                         _u64p]),
    "eon_kzg_commit_ld": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_uint, C.c_size_t, _u64p, _u64p,
                                    C.POINTER(C.c_uint64)]),
    "eon_kzg_commit_lde_ld": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_uint, C.c_size_t, _u64p, _u64p,
                                        C.POINTER(C.c_uint64), C.c_uint, _u64p, _u64p, C.c_size_t]),
    "eon_kzg_evals_on_coset_ld": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint, _u64p, _u64p, C.c_size_t]),
    "eon_msm_srs_range_partial_dev": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, _u64p]),
    "eon_g1_sum_cols_dev": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_size_t, _u64p]),
    "eon_srs_set_range_tables": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint]),
    # multi-device context (one process, several GPUs): same argument lists as the eon_* forms
    "eon_mctx_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "eon_mctx_destroy": (None, [C.c_void_p]),
    "eon_mctx_last_error": (C.c_char_p, [C.c_void_p]),
    "eon_mctx_device_count": (C.c_int, [C.c_void_p]),
    "eon_mctx_ctx": (C.c_void_p, [C.c_void_p, C.c_int]),
    "eon_mctx_launch_count": (C.c_uint64, [C.c_void_p]),
    "eon_mctx_srs_generate_unsafe": (C.c_int, [C.c_void_p, _u64p, C.c_size_t]),
    "eon_mctx_srs_load_affine": (C.c_int, [C.c_void_p, _u64p, C.c_size_t]),
    "eon_mctx_srs_load_compressed": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]),
    "eon_mctx_srs_size": (C.c_size_t, [C.c_void_p]),
    "eon_mctx_srs_read": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, _u64p]),
    "eon_mctx_dft_batch": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t]),
    "eon_mctx_coset_dft_batch": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t, _u64p]),
    "eon_mctx_idft_batch": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t]),
    "eon_mctx_coset_idft_batch": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t, _u64p]),
    "eon_mctx_coset_lde_batch": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_uint, C.c_size_t, C.c_uint, _u64p]),
    "eon_mctx_kzg_commit": (C.c_int, [C.c_void_p, _u64p, C.c_uint, C.c_size_t, _u64p, _u64p, C.POINTER(C.c_uint64)]),
    "eon_mctx_kzg_commit_lde": (C.c_int, [C.c_void_p, _u64p, C.c_uint, C.c_size_t, _u64p, _u64p, C.POINTER(C.c_uint64),
                                          C.c_uint, _u64p, _u64p]),
    "eon_mctx_kzg_commit_coeffs": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_size_t, _u64p, C.POINTER(C.c_uint64)]),
    "eon_mctx_kzg_commit_quotient": (C.c_int, [C.c_void_p, _u64p, C.c_uint, C.c_size_t, C.c_uint, _u64p, _u64p, _u64p]),
    "eon_mctx_kzg_evals_on_coset": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint, _u64p, _u64p]),
    "eon_mctx_kzg_open_batch": (C.c_int, [C.c_void_p, C.c_size_t, _u64p, _u64p, _u64p, _u64p, _u64p]),
    "eon_mctx_kzg_read_coeffs": (C.c_int, [C.c_void_p, C.c_uint64, _u64p]),
    "eon_mctx_handle_dims": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint), C.POINTER(C.c_size_t)]),
    "eon_mctx_handle_free": (C.c_int, [C.c_void_p, C.c_uint64]),
    "eon_mctx_msm_srs": (C.c_int, [C.c_void_p, _u64p, C.c_size_t, C.c_size_t, C.c_size_t, _u64p]),
    "eon_mctx_msm_points": (C.c_int, [C.c_void_p, _u64p, _u64p, C.c_size_t, _u64p]),
    "eon_bench_imad_peak": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "eon_bench_modmul": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "eon_bench_modmul_variant": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "eon_ntt_twiddle_form": (C.c_int, []),
    "eon_last_phase_ms": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "eon_phase_reset": (C.c_int, [C.c_void_p]),
    "eon_phase_name": (C.c_char_p, [C.c_int]),
    "eon_phase_count": (C.c_int, []),
    "eon_bench_copy2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                   C.POINTER(C.c_float)]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)
_lib = None


class EonError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"eon_kzg error {code}: {msg}")
        self.code = code


class DegreeTooLarge(EonError):
    """KzgError::DegreeTooLarge (kzg/src/params.rs:164-173); the reference's prover unwraps it
    into a panic (kzg/src/pcs.rs:238-240)."""


def load():
    """Load libeon_kzg.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class InvalidG1Point(EonError):
    """serde's "Invalid G1 point" (bn254/src/curve.rs:94-96): a compressed encoding that is not a curve point.
    `index` = position of the first such encoding in the batch."""

    def __init__(self, code, msg, index=None):
        super().__init__(code, msg)
        self.index = index


class _PinnedPool:
    """Page-locked host matrices (eon_host_alloc), recycled by size: cudaHostAlloc costs ~0.2 s per GiB, and
    a prover asks for the same shapes over and over."""

    def __init__(self):
        self.free = {}

    def empty(self, shape, dtype=np.uint64):
        import weakref
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        lst = self.free.get(nbytes)
        if lst:
            addr = lst.pop()
        else:
            p = C.c_void_p()
            rc = load().eon_host_alloc(nbytes, C.byref(p))
            if rc != EON_OK:
                raise EonError(rc, "pinned host allocation failed")
            addr = p.value
        buf = (C.c_char * max(nbytes, 1)).from_address(addr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        weakref.finalize(buf, self._recycle, nbytes, addr)   # back to the pool when the last view dies
        return arr

    def _recycle(self, nbytes, addr):
        self.free.setdefault(nbytes, []).append(addr)


_pinned = _PinnedPool()


def pinned_empty(shape, dtype=np.uint64):
    """Uninitialised page-locked numpy array (recycled through a pool when garbage-collected)."""
    return _pinned.empty(shape, dtype)


def ptr(x):
    """Raw address of a numpy array / int / None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"], "buffer must be C-contiguous"
        return x.ctypes.data
    return int(x)


class Context:
    """One eon_ctx (one GPU).  Mirrors the process-global state the reference keeps inside
    Radix2Dit (twiddle cache) and KzgPcs (SRS)."""

    def __init__(self, device=0, stream=None):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.eon_ctx_create(int(device), C.c_void_p(stream or 0), C.byref(h))
        if rc != EON_OK or not h.value:
            raise EonError(rc, "eon_ctx_create failed: no usable sm_100 CUDA device (there is no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.eon_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc == EON_OK:
            return
        msg = self.lib.eon_last_error(self.h).decode()
        if rc == EON_ERR_SRS_TOO_SHORT:
            raise DegreeTooLarge(rc, msg)
        if rc == EON_ERR_BAD_POINT:
            raise InvalidG1Point(rc, msg)
        raise EonError(rc, msg)

    def call(self, name, *args):
        """Call an ABI function.  numpy arrays may be passed directly: they are converted to raw
        addresses here and kept alive (via `args`) until the call returns."""
        conv = []
        for a in args:
            if isinstance(a, np.ndarray):
                assert a.flags["C_CONTIGUOUS"], "buffer must be C-contiguous"
                conv.append(C.c_void_p(a.ctypes.data))
            else:
                conv.append(a)
        self.check(getattr(self.lib, name)(self.h, *conv))

    # -- tiny conveniences -----------------------------------------------------------------
    def sync(self):
        self.call("eon_ctx_sync")

    def launch_count(self):
        return int(self.lib.eon_ctx_launch_count(self.h))

    def srs_size(self):
        return int(self.lib.eon_srs_size(self.h))

    # -- compressed G1 points (bn254/src/curve.rs:84-98,136-139) ---------------------------------
    def g1_to_bytes(self, points_wire, enc=G1_ENC_HALO2):
        """G1::to_bytes for a batch: uint64 [n, 8] affine wire points -> uint8 [n, 32]."""
        p = np.ascontiguousarray(points_wire, dtype=np.uint64).reshape(-1, 8)
        out = np.zeros((p.shape[0], 32), dtype=np.uint8)
        self.call("eon_g1_compress", p, p.shape[0], out, int(enc))
        return out

    def g1_from_bytes(self, data, enc=G1_ENC_HALO2):
        """Deserialize for G1 for a batch: uint8 [n, 32] -> uint64 [n, 8]; raises InvalidG1Point."""
        b = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 32)
        out = np.zeros((b.shape[0], 8), dtype=np.uint64)
        bad = C.c_size_t(0)
        try:
            self.call("eon_g1_decompress", b, b.shape[0], out, int(enc), C.byref(bad))
        except InvalidG1Point as e:
            e.index = int(bad.value)
            raise
        return out

    def srs_load_compressed(self, data, enc=G1_ENC_HALO2):
        """g1_powers of a deserialised StructuredReferenceString (kzg/src/params.rs:56-77) as 32-byte
        compressed points, decompressed on the device into the resident SRS."""
        b = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 32)
        bad = C.c_size_t(0)
        try:
            self.call("eon_srs_load_compressed", b, b.shape[0], int(enc), C.byref(bad))
        except InvalidG1Point as e:
            e.index = int(bad.value)
            raise

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self.call("eon_dev_alloc", nbytes, C.byref(p))
        return p.value

    def dev_free(self, p):
        self.call("eon_dev_free", C.c_void_p(p))

    def h2d(self, dptr, arr):
        self.call("eon_h2d", C.c_void_p(dptr), arr, arr.nbytes)

    def d2h(self, arr, dptr):
        self.call("eon_d2h", arr, C.c_void_p(dptr), arr.nbytes)

    def phase_reset(self):
        self.call("eon_phase_reset")

    def phase_ms(self):
        out = {}
        v = C.c_float()
        for ph in range(int(self.lib.eon_phase_count())):
            self.call("eon_last_phase_ms", ph, C.byref(v))
            out[self.lib.eon_phase_name(ph).decode()] = float(v.value)
        return out

    def imad_peak_tops(self, kind=0):
        v = C.c_double()
        self.call("eon_bench_imad_peak", kind, C.byref(v))
        return float(v.value)

    def modmul_gmuls(self, field=1, variant=0):
        """1e9 Montgomery products/s; variant 0 = the library's product, 1 = word-serial, 2 = split, 3 = square."""
        v = C.c_double()
        self.call("eon_bench_modmul_variant", field, variant, C.byref(v))
        return float(v.value)


class _ShardContext(Context):
    """Non-owning view of one device context of a MultiContext (tuning / timing calls)."""

    def __init__(self, lib, handle, device):  # noqa: D401 - no eon_ctx_create here
        self.lib = lib
        self.h = C.c_void_p(handle)
        self.device = device

    def close(self):
        self.h = C.c_void_p()


class MultiContext:
    """One eon_mctx: several GPUs behind the whole-matrix calls of one process (include/eon_kzg.h).  Quacks like
    Context for the host mirrors (GpuDft, GpuKzgPcs, KzgMmcs): `call("eon_X", ...)` goes to `eon_mctx_X` when that
    exists; calls without per-device state (point compression, sums) go to the first device."""

    _FIRST_DEVICE = {"eon_g1_compress", "eon_g1_decompress", "eon_g1_sum", "eon_bench_imad_peak",
                     "eon_bench_modmul", "eon_bench_modmul_variant"}

    def __init__(self, devices):
        self.lib = load()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self.lib.eon_mctx_create(devs, len(devices), C.byref(h))
        if rc != EON_OK or not h.value:
            raise EonError(rc, "eon_mctx_create failed: no usable sm_100 CUDA device (there is no CPU fallback)")
        self.h = h
        self.devices = list(devices)
        self.device = devices[0]

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.eon_mctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard(self, i):
        p = self.lib.eon_mctx_ctx(self.h, int(i))
        if not p:
            raise IndexError(i)
        return _ShardContext(self.lib, p, self.devices[i])

    def _raise(self, rc, msg):
        if rc == EON_ERR_SRS_TOO_SHORT:
            raise DegreeTooLarge(rc, msg)
        if rc == EON_ERR_BAD_POINT:
            raise InvalidG1Point(rc, msg)
        raise EonError(rc, msg)

    def call(self, name, *args):
        conv = []
        for a in args:
            if isinstance(a, np.ndarray):
                assert a.flags["C_CONTIGUOUS"], "buffer must be C-contiguous"
                conv.append(C.c_void_p(a.ctypes.data))
            else:
                conv.append(a)
        mname = "eon_mctx_" + name[4:]
        if mname in _SIGS:
            rc = getattr(self.lib, mname)(self.h, *conv)
            if rc != EON_OK:
                self._raise(rc, self.lib.eon_mctx_last_error(self.h).decode())
            return
        if name == "eon_kzg_open":  # one matrix = a batch of one
            handle, pts, npoints, vals, wits = conv
            hs = np.array([handle.value if hasattr(handle, "value") else int(handle)], dtype=np.uint64)
            npts = np.array([int(npoints)], dtype=np.uint64)
            return self.call("eon_kzg_open_batch", 1, hs, npts, pts, vals, wits)
        if name in self._FIRST_DEVICE:
            return self.shard(0).call(name, *args)
        raise EonError(EON_ERR_BAD_ARG, f"{name} has no multi-device form (use MultiContext.shard(i))")

    def sync(self):
        for i in range(len(self.devices)):
            self.shard(i).sync()

    def launch_count(self):
        return int(self.lib.eon_mctx_launch_count(self.h))

    def srs_size(self):
        return int(self.lib.eon_mctx_srs_size(self.h))

    def g1_to_bytes(self, points_wire, enc=G1_ENC_HALO2):
        return self.shard(0).g1_to_bytes(points_wire, enc)

    def g1_from_bytes(self, data, enc=G1_ENC_HALO2):
        return self.shard(0).g1_from_bytes(data, enc)

    def srs_load_compressed(self, data, enc=G1_ENC_HALO2):
        b = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 32)
        bad = C.c_size_t(0)
        try:
            self.call("eon_srs_load_compressed", b, b.shape[0], int(enc), C.byref(bad))
        except InvalidG1Point as e:
            e.index = int(bad.value)
            raise


_default_ctx = {}


def default_context(device=0):
    """Lazily-created per-device context (TwoAdicSubgroupDft: Clone + Default needs a
    process-global, SURVEY §7 hard part 7)."""
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]

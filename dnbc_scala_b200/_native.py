"""ctypes binding of libdnbc_b200.so (the C ABI in include/dnbc_b200.h).

This is the same boundary the reference-side JNI glue binds; Python is
used here only because the container has no JVM.  There is no CPU fallback: if the shared
library is missing, or no sm_100 GPU is present, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DNBC_B200_LIB") or os.path.join(_HERE, "lib", "libdnbc_b200.so")   # override: A/B builds only
CSRC = os.path.join(_HERE, "csrc")

OK, EINVAL, ECUDA, ENODEVICE, ESTATE, ENOMEM, ECOMM = range(7)
COMM_ID_BYTES = 128
OPT_MAX_CHUNK_POINTS, OPT_FB2_INTERLEAVE, OPT_FUSED_ESTEP, OPT_FB_SEQ_LANES = 1, 2, 3, 4
DECODE_REFERENCE, DECODE_BACKTRACE = 0, 1
PREC_F32, PREC_F64 = 0, 1
QUIRK_TRANSITIONS, DOUBLE_COUNT_FIRST = 1, 2
FAMILIES = ("emission", "forward", "backward", "xi", "stats_gmm", "stats_cat", "mstep", "decode")

# every symbol include/dnbc_b200.h declares (tests check the library exports all of them)
SYMBOLS = (
    "dnbc_abi_version", "dnbc_create", "dnbc_destroy", "dnbc_last_error", "dnbc_set_params",
    "dnbc_get_params", "dnbc_upload", "dnbc_generate", "dnbc_download", "dnbc_download_f32", "dnbc_data_shape",
    "dnbc_emission_logprob", "dnbc_supervised_counts", "dnbc_mle", "dnbc_stats_size",
    "dnbc_em_estep", "dnbc_em_estep_streamed", "dnbc_em_stats_device_ptr", "dnbc_em_get_stats", "dnbc_em_set_stats",
    "dnbc_em_mstep", "dnbc_em_set_variance_floor", "dnbc_em_floored_components", "dnbc_em_iterate",
    "dnbc_em_get_posteriors", "dnbc_comm_unique_id", "dnbc_comm_init_rank", "dnbc_comm_init_all", "dnbc_comm_info",
    "dnbc_comm_destroy", "dnbc_em_allreduce", "dnbc_em_allreduce_group", "dnbc_em_iterate_sharded",
    "dnbc_em_iterate_group", "dnbc_decode",
    "dnbc_set_option", "dnbc_synchronize", "dnbc_stream", "dnbc_launch_count", "dnbc_profile_enable",
    "dnbc_profile_read",
)


class DnbcError(Exception):
    """Mirrors the reference's plain `java.lang.Exception(message)` convention."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def build(force: bool = False) -> str:
    """Compile the CUDA library for sm_100a with the committed Makefile (nvcc)."""
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", CSRC, "-j8"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


class _Config(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("n_disc", C.c_int32), ("n_cont", C.c_int32),
                ("alphabet", C.c_void_p), ("mixture", C.c_void_p), ("device", C.c_int32),
                ("flags", C.c_int32)]


_lib = None


def _preload_nccl():
    """libdnbc_b200.so links libnccl.so.2.  Inside a Python process that also imports torch, load
    torch's bundled NCCL first so that both resolve the SONAME to the same library (a second,
    older copy pulled in ahead of torch could miss symbols torch needs)."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for base in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
        path = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(path):
            try:
                C.CDLL(path, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            return


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(dnbc_scala_b200 has no CPU fallback)")
        _preload_nccl()
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.dnbc_abi_version.restype = i32
        L.dnbc_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
        L.dnbc_destroy.argtypes = [vp]
        L.dnbc_last_error.argtypes = [vp]
        L.dnbc_last_error.restype = C.c_char_p
        L.dnbc_set_params.argtypes = [vp] * 8
        L.dnbc_get_params.argtypes = [vp] * 8
        L.dnbc_upload.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        L.dnbc_generate.argtypes = [vp, i64, C.c_int32, C.c_uint64]
        L.dnbc_download.argtypes = [vp, i64, i64, vp, vp, vp, vp]
        L.dnbc_download_f32.argtypes = [vp, i64, i64, vp, vp]
        L.dnbc_data_shape.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
        L.dnbc_emission_logprob.argtypes = [vp, i32, vp]
        L.dnbc_supervised_counts.argtypes = [vp, i32, vp, vp, vp]
        L.dnbc_mle.argtypes = [vp, i32, vp, i32, dbl, vp, vp]
        L.dnbc_stats_size.argtypes = [vp]
        L.dnbc_stats_size.restype = i64
        L.dnbc_em_estep.argtypes = [vp, i32]
        L.dnbc_em_estep_streamed.argtypes = [vp, i64, vp, vp, vp, vp, i32]
        L.dnbc_em_stats_device_ptr.argtypes = [vp, C.POINTER(vp)]
        L.dnbc_em_get_stats.argtypes = [vp, vp]
        L.dnbc_em_set_stats.argtypes = [vp, vp]
        L.dnbc_em_mstep.argtypes = [vp, i64]
        L.dnbc_em_iterate.argtypes = [vp, i32, vp]
        L.dnbc_em_set_variance_floor.argtypes = [vp, dbl]
        L.dnbc_em_floored_components.argtypes = [vp, C.POINTER(i64)]
        L.dnbc_comm_unique_id.argtypes = [vp]
        L.dnbc_comm_init_rank.argtypes = [vp, i32, i32, vp]
        L.dnbc_comm_init_all.argtypes = [vp, i32]
        L.dnbc_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
        L.dnbc_comm_destroy.argtypes = [vp]
        L.dnbc_em_allreduce.argtypes = [vp]
        L.dnbc_em_allreduce_group.argtypes = [vp, i32]
        L.dnbc_em_iterate_sharded.argtypes = [vp, i32, i64, vp]
        L.dnbc_em_iterate_group.argtypes = [vp, i32, i32, vp]
        L.dnbc_em_get_posteriors.argtypes = [vp, vp]
        L.dnbc_decode.argtypes = [vp, i32, i32, vp, vp, vp]
        L.dnbc_set_option.argtypes = [vp, i32, i64]
        L.dnbc_synchronize.argtypes = [vp]
        L.dnbc_stream.argtypes = [vp, C.POINTER(vp)]
        L.dnbc_launch_count.argtypes = [vp]
        L.dnbc_launch_count.restype = i64
        L.dnbc_profile_enable.argtypes = [vp, i32]
        L.dnbc_profile_read.argtypes = [vp, vp, vp, i32]
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class Context:
    """One `dnbc_ctx`: a model shape bound to one GPU, its parameters and resident data."""

    def __init__(self, n_states: int, alphabet: Sequence[int], mixture: Sequence[int], device: int = 0):
        self._h = None
        self.device = int(device)
        self.N = int(n_states)
        self.V = np.ascontiguousarray(alphabet, np.int32).reshape(-1)
        self.K = np.ascontiguousarray(mixture, np.int32).reshape(-1)
        self.D, self.C = len(self.V), len(self.K)
        self.Vmax = int(max([1] + [int(v) for v in self.V]))
        self.Kmax = int(max([1] + [int(k) for k in self.K]))
        cfg = _Config(self.N, self.D, self.C, _p(self.V) if self.D else None,
                      _p(self.K) if self.C else None, int(device), 0)
        h = C.c_void_p()
        rc = lib().dnbc_create(C.byref(cfg), C.byref(h))
        if rc != OK:
            raise DnbcError(rc, (lib().dnbc_last_error(None) or b"").decode())
        self._h = h
        self.S = self.P = 0

    # -- plumbing -----------------------------------------------------------------
    def _check(self, rc: int):
        if rc != OK:
            raise DnbcError(rc, (lib().dnbc_last_error(self._h) or b"").decode())

    def close(self):
        if self._h is not None:
            lib().dnbc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- parameters ---------------------------------------------------------------
    def set_params(self, pi, A, B, w, mu, var, active=None):
        f = lambda a, shape: np.ascontiguousarray(a, np.float64).reshape(shape)
        pi, A = f(pi, (self.N,)), f(A, (self.N, self.N))
        B = f(B, (self.D, self.N, self.Vmax)) if self.D else None
        g = (self.C, self.N, self.Kmax)
        w, mu, var = (f(w, g), f(mu, g), f(var, g)) if self.C else (None, None, None)
        act = None if active is None else np.ascontiguousarray(active, np.uint8).reshape(self.N)
        self._check(lib().dnbc_set_params(self._h, _p(pi), _p(A), _p(B), _p(w), _p(mu), _p(var), _p(act)))

    def get_params(self):
        pi = np.zeros(self.N)
        A = np.zeros((self.N, self.N))
        B = np.zeros((self.D, self.N, self.Vmax))
        w = np.zeros((self.C, self.N, self.Kmax))
        mu, var = np.zeros_like(w), np.zeros_like(w)
        active = np.zeros(self.N, np.uint8)
        self._check(lib().dnbc_get_params(self._h, _p(pi), _p(A), _p(B), _p(w), _p(mu), _p(var), _p(active)))
        return dict(pi=pi, A=A, B=B, w=w, mu=mu, var=var, active=active)

    # -- observations -------------------------------------------------------------
    def upload(self, offsets, disc, cont, labels=None, cont_dtype=np.float64):
        off = np.ascontiguousarray(offsets, np.int64)
        P = int(off[-1])
        d = np.ascontiguousarray(disc, np.uint8).reshape(self.D, P) if self.D else None
        c = np.ascontiguousarray(cont, cont_dtype).reshape(self.C, P) if self.C else None
        lab = None if labels is None else np.ascontiguousarray(labels, np.int32).reshape(P)
        c64 = c if (c is not None and c.dtype == np.float64) else None
        c32 = c if (c is not None and c.dtype == np.float32) else None
        self._check(lib().dnbc_upload(self._h, len(off) - 1, _p(off), _p(d), _p(c64), _p(c32), _p(lab)))
        self.S, self.P = len(off) - 1, P

    def generate(self, n_seq: int, seq_len: int, seed: int):
        self._check(lib().dnbc_generate(self._h, int(n_seq), int(seq_len), int(seed)))
        self.S, self.P = int(n_seq), int(n_seq) * int(seq_len)

    def download(self, first: int = 0, count: Optional[int] = None, labels: bool = True):
        count = self.S - first if count is None else count
        off = np.zeros(count + 1, np.int64)
        self._check(lib().dnbc_download(self._h, first, count, _p(off), None, None, None))
        P = int(off[-1])
        disc = np.zeros((self.D, P), np.uint8)
        cont = np.zeros((self.C, P), np.float64)
        lab = np.zeros(P, np.int32) if labels else None
        self._check(lib().dnbc_download(self._h, first, count, _p(off), _p(disc), _p(cont), _p(lab)))
        return off, disc, cont, lab

    def upload_raw(self, n_seq: int, offsets: np.ndarray, disc_ptr: int, cont32_ptr: int, labels_ptr: int = 0):
        """Upload from caller-owned (e.g. pinned) host buffers given as raw addresses."""
        off = np.ascontiguousarray(offsets, np.int64)
        self._check(lib().dnbc_upload(self._h, int(n_seq), _p(off), disc_ptr or None, None, cont32_ptr or None,
                                      labels_ptr or None))
        self.S, self.P = int(n_seq), int(off[-1])

    def download_f32_raw(self, first: int, count: int, disc_ptr: int, cont32_ptr: int):
        self._check(lib().dnbc_download_f32(self._h, first, count, disc_ptr or None, cont32_ptr or None))

    # -- kernels ------------------------------------------------------------------
    def emission_logprob(self, precision: int = PREC_F32) -> np.ndarray:
        out = np.zeros((self.P, self.N), np.float64 if precision == PREC_F64 else np.float32)
        self._check(lib().dnbc_emission_logprob(self._h, precision, _p(out)))
        return out

    def supervised_counts(self, flags: int = 0):
        cnt_pi = np.zeros(self.N, np.int64)
        cnt_A = np.zeros((self.N, self.N), np.int64)
        cnt_B = np.zeros((self.D, self.N, self.Vmax), np.int64)
        self._check(lib().dnbc_supervised_counts(self._h, flags, _p(cnt_pi), _p(cnt_A), _p(cnt_B)))
        return cnt_pi, cnt_A, cnt_B

    def mle(self, flags: int = 0, init_centroids=None, max_iters: int = 30, rel_tol: float = 0.01):
        cen = None
        if init_centroids is not None:
            cen = np.ascontiguousarray(init_centroids, np.float64).reshape(self.C, self.N, self.Kmax)
        ll = np.zeros((self.C, self.N, max_iters + 1))
        its = np.zeros((self.C, self.N), np.int32)
        self._check(lib().dnbc_mle(self._h, flags, _p(cen), max_iters, rel_tol, _p(ll), _p(its)))
        return ll, its

    def stats_size(self) -> int:
        return int(lib().dnbc_stats_size(self._h))

    def em_estep(self, clamp_labels: bool = False):
        self._check(lib().dnbc_em_estep(self._h, int(clamp_labels)))

    def em_estep_streamed_raw(self, n_seq: int, offsets: np.ndarray, disc_ptr: int, cont32_ptr: int,
                              labels_ptr: int = 0, clamp_labels: bool = False):
        """Upload (from raw, ideally pinned, host addresses) overlapped with the E-step."""
        off = np.ascontiguousarray(offsets, np.int64)
        self._check(lib().dnbc_em_estep_streamed(self._h, int(n_seq), _p(off), disc_ptr or None, cont32_ptr or None,
                                                 labels_ptr or None, int(clamp_labels)))
        self.S, self.P = int(n_seq), int(off[-1])

    def em_estep_streamed(self, offsets, disc, cont32, labels=None, clamp_labels: bool = False):
        off = np.ascontiguousarray(offsets, np.int64)
        P = int(off[-1])
        d = np.ascontiguousarray(disc, np.uint8).reshape(self.D, P) if self.D else None
        c = np.ascontiguousarray(cont32, np.float32).reshape(self.C, P) if self.C else None
        lab = None if labels is None else np.ascontiguousarray(labels, np.int32).reshape(P)
        self._check(lib().dnbc_em_estep_streamed(self._h, len(off) - 1, _p(off), _p(d), _p(c), _p(lab), int(clamp_labels)))
        self.synchronize()                     # numpy sources are pageable: finish before they can be freed
        self.S, self.P = len(off) - 1, P

    def em_stats_device_ptr(self) -> int:
        p = C.c_void_p()
        self._check(lib().dnbc_em_stats_device_ptr(self._h, C.byref(p)))
        return int(p.value)

    def em_get_stats(self) -> np.ndarray:
        st = np.zeros(self.stats_size())
        self._check(lib().dnbc_em_get_stats(self._h, _p(st)))
        return st

    def em_set_stats(self, st):
        st = np.ascontiguousarray(st, np.float64)
        assert st.size == self.stats_size()
        self._check(lib().dnbc_em_set_stats(self._h, _p(st)))

    def em_mstep(self, total_sequences: int):
        self._check(lib().dnbc_em_mstep(self._h, int(total_sequences)))

    def em_iterate(self, n_iters: int) -> np.ndarray:
        ll = np.zeros(n_iters)
        self._check(lib().dnbc_em_iterate(self._h, n_iters, _p(ll)))
        return ll

    def em_set_variance_floor(self, rel_floor: float):
        self._check(lib().dnbc_em_set_variance_floor(self._h, float(rel_floor)))

    def em_floored_components(self) -> int:
        n = C.c_int64(0)
        self._check(lib().dnbc_em_floored_components(self._h, C.byref(n)))
        return int(n.value)

    # -- multi-GPU (the library owns the NCCL communicator) ---------------------------
    def comm_init_rank(self, n_ranks: int, rank: int, unique_id: bytes):
        assert len(unique_id) == COMM_ID_BYTES
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._check(lib().dnbc_comm_init_rank(self._h, int(n_ranks), int(rank), C.cast(buf, C.c_void_p)))

    def comm_info(self):
        n, r = C.c_int(0), C.c_int(0)
        self._check(lib().dnbc_comm_info(self._h, C.byref(n), C.byref(r)))
        return int(n.value), int(r.value)

    def comm_destroy(self):
        self._check(lib().dnbc_comm_destroy(self._h))

    def em_allreduce(self):
        self._check(lib().dnbc_em_allreduce(self._h))

    def em_iterate_sharded(self, n_iters: int, total_sequences: int) -> np.ndarray:
        ll = np.zeros(n_iters)
        self._check(lib().dnbc_em_iterate_sharded(self._h, n_iters, int(total_sequences), _p(ll)))
        return ll

    def em_posteriors(self) -> np.ndarray:
        g = np.zeros((self.P, self.N), np.float32)
        self._check(lib().dnbc_em_get_posteriors(self._h, _p(g)))
        return g

    def decode(self, mode: int = DECODE_REFERENCE, precision: int = PREC_F64, want_path=True, want_score=True,
               want_tie=True):
        path = np.zeros(self.P, np.uint16) if want_path else None
        score = np.zeros(self.S) if want_score else None
        tie = np.zeros(self.P, np.uint8) if want_tie else None
        self._check(lib().dnbc_decode(self._h, mode, precision, _p(path), _p(score), _p(tie)))
        return path, score, tie

    def decode_raw(self, mode: int, precision: int, path_ptr: int = 0, score_ptr: int = 0, tie_ptr: int = 0):
        """dnbc_decode into caller-owned host buffers (u16 [P], f64 [S], u8 [P]; 0 = not wanted).
        Pinned buffers make the result copies asynchronous DMA transfers."""
        self._check(lib().dnbc_decode(self._h, mode, precision, C.c_void_p(path_ptr or None),
                                      C.c_void_p(score_ptr or None), C.c_void_p(tie_ptr or None)))

    def set_option(self, option: int, value: int):
        self._check(lib().dnbc_set_option(self._h, int(option), int(value)))

    def synchronize(self):
        self._check(lib().dnbc_synchronize(self._h))

    def stream(self) -> int:
        p = C.c_void_p()
        self._check(lib().dnbc_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def launch_count(self) -> int:
        return int(lib().dnbc_launch_count(self._h))

    def profile_enable(self, on: bool):
        self._check(lib().dnbc_profile_enable(self._h, int(on)))

    def profile_read(self, reset: bool = True):
        ms = np.zeros(len(FAMILIES))
        n = np.zeros(len(FAMILIES), np.int64)
        self._check(lib().dnbc_profile_read(self._h, _p(ms), _p(n), int(reset)))
        return dict(zip(FAMILIES, ms.tolist())), dict(zip(FAMILIES, n.tolist()))


def comm_unique_id() -> bytes:
    """128-byte NCCL unique id; rank 0 creates it and ships it to the other ranks."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = lib().dnbc_comm_unique_id(C.cast(buf, C.c_void_p))
    if rc != OK:
        raise DnbcError(rc, (lib().dnbc_last_error(None) or b"").decode())
    return buf.raw


def _handles(ctxs):
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    return arr


def comm_init_all(ctxs):
    """One process driving several GPUs: join the contexts (one per device) in one communicator."""
    arr = _handles(ctxs)
    rc = lib().dnbc_comm_init_all(C.cast(arr, C.c_void_p), len(ctxs))
    if rc != OK:
        raise DnbcError(rc, (lib().dnbc_last_error(ctxs[0]._h) or b"").decode())


def em_allreduce_group(ctxs):
    arr = _handles(ctxs)
    rc = lib().dnbc_em_allreduce_group(C.cast(arr, C.c_void_p), len(ctxs))
    if rc != OK:
        raise DnbcError(rc, (lib().dnbc_last_error(ctxs[0]._h) or b"").decode())


def em_iterate_group(ctxs, n_iters: int) -> np.ndarray:
    arr = _handles(ctxs)
    ll = np.zeros(n_iters)
    rc = lib().dnbc_em_iterate_group(C.cast(arr, C.c_void_p), len(ctxs), int(n_iters), _p(ll))
    if rc != OK:
        raise DnbcError(rc, (lib().dnbc_last_error(ctxs[0]._h) or b"").decode())
    return ll

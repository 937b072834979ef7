"""ctypes bindings of include/b2l.h. The product path: fails loudly when libb2l.so is missing
or no CUDA device is present -- there is no CPU fallback and nothing here touches oracle/."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libb2l.so")
LIB_PATH = os.environ.get("B2L_LIB_PATH", LIB_PATH)   # development builds (profiling variants); the default is the in-tree library

# every symbol include/b2l.h declares (tests/test_capi_symbols.py checks the header against this)
SYMBOLS = [
    "b2l_create", "b2l_nccl_unique_id", "b2l_shard_window", "b2l_is_model_tensor", "b2l_upload_tensor", "b2l_synth_tensor", "b2l_finalize", "b2l_destroy",
    "b2l_prefill", "b2l_decode", "b2l_decode_loop", "b2l_get_logits", "b2l_set_taps", "b2l_get_hidden",
    "b2l_get_kv_page", "b2l_get_info", "b2l_set_decode_mode", "b2l_set_prefill_mode", "b2l_debug_mega_profile", "b2l_last_error", "b2l_op_gemv", "b2l_op_argmax", "b2l_op_gemm_bf16",
]


class B2lParams(C.Structure):
    _fields_ = [
        ("hidden_size", C.c_int32), ("intermediate_size", C.c_int32), ("num_layers", C.c_int32),
        ("num_heads", C.c_int32), ("num_kv_heads", C.c_int32), ("head_dim", C.c_int32),
        ("vocab_size", C.c_int32), ("tie_word_embeddings", C.c_int32), ("rms_norm_eps", C.c_float),
        ("max_batch", C.c_int32), ("max_positions", C.c_int32), ("page_size", C.c_int32),
        ("num_pages", C.c_int32), ("max_prefill_tokens", C.c_int32),
        ("tp_rank", C.c_int32), ("tp_size", C.c_int32), ("device", C.c_int32),
    ]


class B2lInfo(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
        ("hbm_bytes", C.c_int64), ("weight_bytes", C.c_int64), ("kv_bytes", C.c_int64),
        ("stream_bytes_per_token", C.c_int64), ("kernels_launched", C.c_int64), ("decode_mode", C.c_int32),
        ("batched_tensor_core", C.c_int32), ("tp_transport", C.c_int32),
        ("device_name", C.c_char * 64),
    ]


class B2lError(RuntimeError):
    pass


_LIB = None


def lib():
    """Load libb2l.so; raise (never fall back) when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise B2lError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                       "gabby_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.b2l_create.argtypes = [C.POINTER(B2lParams), vp, vp, C.POINTER(vp)]
    L.b2l_nccl_unique_id.argtypes = [vp]
    L.b2l_upload_tensor.argtypes = [vp, C.c_char_p, vp, C.POINTER(C.c_int64), C.c_int]
    L.b2l_synth_tensor.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64), C.c_int, C.c_uint32, C.c_float, C.c_float]
    L.b2l_shard_window.argtypes = [C.POINTER(B2lParams), C.c_char_p, C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int64)]
    if hasattr(L, "b2l_is_model_tensor"):   # absent from older development builds selected with B2L_LIB_PATH
        L.b2l_is_model_tensor.argtypes = [C.c_char_p]
    L.b2l_finalize.argtypes = [vp]
    L.b2l_destroy.argtypes = [vp]
    L.b2l_destroy.restype = None
    L.b2l_prefill.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp]
    L.b2l_decode.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp]
    L.b2l_decode_loop.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, vp, C.POINTER(C.c_float)]
    L.b2l_get_logits.argtypes = [vp, C.c_int, C.c_int, vp]
    L.b2l_set_taps.argtypes = [vp, C.c_int]
    L.b2l_get_hidden.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.b2l_get_kv_page.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.b2l_get_info.argtypes = [vp, C.POINTER(B2lInfo)]
    L.b2l_set_decode_mode.argtypes = [vp, C.c_int]
    L.b2l_set_prefill_mode.argtypes = [vp, C.c_int]
    L.b2l_debug_mega_profile.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_int), vp]
    L.b2l_last_error.argtypes = [vp]
    L.b2l_last_error.restype = C.c_char_p
    L.b2l_op_gemv.argtypes = [C.c_int, vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.POINTER(C.c_float)]
    L.b2l_op_argmax.argtypes = [C.c_int, vp, C.c_int, C.c_int, vp]
    L.b2l_op_gemm_bf16.argtypes = [C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
    _LIB = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Engine:
    """Thin object wrapper over the C-ABI (used by tests and bench.py; the production host layer
    is C++: gabby_b200/host/)."""

    def __init__(self, arch, rope_cos_sin: np.ndarray, *, max_batch=1, max_positions=1024, page_size=16,
                 num_pages=None, max_prefill_tokens=1024, device=0, tp_rank=0, tp_size=1, nccl_id: bytes | None = None):
        self.L = lib()
        self.arch = arch
        if num_pages is None:
            num_pages = max_batch * ((max_positions + page_size - 1) // page_size)
        rope = np.ascontiguousarray(rope_cos_sin, dtype=np.float32)
        assert rope.shape == (max_positions, arch.head_dim // 2, 2), rope.shape
        self.params = B2lParams(arch.hidden_size, arch.intermediate_size, arch.num_hidden_layers,
                                arch.num_attention_heads, arch.num_key_value_heads, arch.head_dim, arch.vocab_size,
                                int(arch.tie_word_embeddings), arch.rms_norm_eps, max_batch, max_positions, page_size,
                                num_pages, max_prefill_tokens, tp_rank, tp_size, device)
        h = C.c_void_p()
        idbuf = (C.c_char * 128).from_buffer_copy(nccl_id) if nccl_id else None
        if self.L.b2l_create(C.byref(self.params), _p(rope), idbuf, C.byref(h)) != 0:
            raise B2lError("b2l_create: " + self.L.b2l_last_error(None).decode())
        self.h = h
        self.page_size, self.num_pages, self.max_positions = page_size, num_pages, max_positions
        self.max_blocks = (max_positions + page_size - 1) // page_size

    def _ck(self, rc, what):
        if rc != 0:
            raise B2lError(f"{what}: " + self.L.b2l_last_error(self.h).decode())

    def upload(self, name: str, bits: np.ndarray, shape):
        a = np.ascontiguousarray(bits)
        assert a.dtype == np.uint16
        sh = (C.c_int64 * len(shape))(*shape)
        self._ck(self.L.b2l_upload_tensor(self.h, name.encode(), _p(a), sh, len(shape)), "upload " + name)

    def synth(self, name: str, shape, tensor_seed: int, scale: float, offset: float):
        sh = (C.c_int64 * len(shape))(*shape)
        self._ck(self.L.b2l_synth_tensor(self.h, name.encode(), sh, len(shape), tensor_seed, scale, offset), "synth " + name)

    def finalize(self):
        self._ck(self.L.b2l_finalize(self.h), "finalize")

    def _bt(self, block_tables):
        bt = _i32(block_tables)
        if bt.ndim == 1:
            bt = bt[None, :]
        return bt, bt.shape[1]

    def prefill(self, tokens_per_seq, ctx_lens, block_tables) -> np.ndarray:
        toks = _i32(np.concatenate([np.asarray(t, dtype=np.int32) for t in tokens_per_seq]))
        q_lens = _i32([len(t) for t in tokens_per_seq])
        ctx = _i32(ctx_lens)
        bt, mb = self._bt(block_tables)
        out = np.empty(len(q_lens), dtype=np.int32)
        self._ck(self.L.b2l_prefill(self.h, len(q_lens), _p(toks), _p(q_lens), _p(ctx), _p(bt), mb, _p(out)), "prefill")
        return out

    def decode(self, tokens, positions, block_tables) -> np.ndarray:
        t, p = _i32(tokens), _i32(positions)
        bt, mb = self._bt(block_tables)
        out = np.empty(t.size, dtype=np.int32)
        self._ck(self.L.b2l_decode(self.h, t.size, _p(t), _p(p), _p(bt), mb, _p(out)), "decode")
        return out

    def decode_loop(self, tokens, positions, block_tables, n_steps: int):
        t, p = _i32(tokens), _i32(positions)
        bt, mb = self._bt(block_tables)
        out = np.empty((n_steps, t.size), dtype=np.int32)
        ms = C.c_float(0)
        self._ck(self.L.b2l_decode_loop(self.h, t.size, _p(t), _p(p), _p(bt), mb, n_steps, _p(out), C.byref(ms)), "decode_loop")
        return out, ms.value

    def logits(self, row0=0, n_rows=1) -> np.ndarray:
        out = np.empty((n_rows, self.arch.vocab_size), dtype=np.float32)
        self._ck(self.L.b2l_get_logits(self.h, row0, n_rows, _p(out)), "get_logits")
        return out

    def set_taps(self, on: bool):
        self._ck(self.L.b2l_set_taps(self.h, int(on)), "set_taps")

    def hidden(self, slab: int, row0: int, n_rows: int) -> np.ndarray:
        out = np.empty((n_rows, self.arch.hidden_size), dtype=np.float32)
        self._ck(self.L.b2l_get_hidden(self.h, slab, row0, n_rows, _p(out)), "get_hidden")
        return out

    def kv_page(self, layer: int, page: int, which: int) -> np.ndarray:
        kvd = self.arch.num_key_value_heads // self.params.tp_size * self.arch.head_dim
        out = np.empty((self.page_size, kvd), dtype=np.uint16)
        self._ck(self.L.b2l_get_kv_page(self.h, layer, page, which, _p(out)), "get_kv_page")
        return out

    def info(self) -> B2lInfo:
        i = B2lInfo()
        self._ck(self.L.b2l_get_info(self.h, C.byref(i)), "get_info")
        return i

    def mega_profile(self, enable=True):
        """-> (ns [16 + 2*160][n_phases+1] uint64, phase_types [n_phases]) of the last megakernel token: rows 0-15 summary
        (CTA 0 / last CTA), rows 16.. input-ready time per CTA, rows 176.. phase-end time per CTA"""
        n = C.c_int(0)
        self._ck(self.L.b2l_debug_mega_profile(self.h, int(enable), None, C.byref(n), None), "mega_profile")
        ns = np.zeros((16 + 2 * 160 + 64, n.value + 1), dtype=np.uint64)   # kMegaProfRows
        types = np.zeros(n.value, dtype=np.int32)
        self._ck(self.L.b2l_debug_mega_profile(self.h, int(enable), _p(ns), C.byref(n), _p(types)), "mega_profile")
        return ns, types

    def set_prefill_mode(self, mode: int):
        self._ck(self.L.b2l_set_prefill_mode(self.h, mode), "set_prefill_mode")

    def set_decode_mode(self, mode: int):
        self._ck(self.L.b2l_set_decode_mode(self.h, mode), "set_decode_mode")

    def close(self):
        if getattr(self, "h", None):
            self.L.b2l_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_window(arch, name: str, shape, tp_rank: int, tp_size: int):
    """(row0, nrows, col0, ncols) of the full HF tensor that `tp_rank` keeps. No GPU needed."""
    p = B2lParams(arch.hidden_size, arch.intermediate_size, arch.num_hidden_layers, arch.num_attention_heads,
                  arch.num_key_value_heads, arch.head_dim, arch.vocab_size, int(arch.tie_word_embeddings), arch.rms_norm_eps,
                  1, 64, 16, 4, 64, tp_rank, tp_size, 0)
    sh = (C.c_int64 * len(shape))(*shape)
    win = (C.c_int64 * 4)()
    if lib().b2l_shard_window(C.byref(p), name.encode(), sh, len(shape), win) != 0:
        raise B2lError("b2l_shard_window: " + lib().b2l_last_error(None).decode())
    return tuple(win)


def nccl_unique_id() -> bytes:
    """128-byte NCCL unique id (rank 0 makes it, every TP rank passes it to Engine(nccl_id=...))."""
    buf = C.create_string_buffer(128)
    if lib().b2l_nccl_unique_id(buf) != 0:
        raise B2lError("b2l_nccl_unique_id: " + lib().b2l_last_error(None).decode())
    return buf.raw


def op_gemv(W_bits, x, y_in=None, norm_w_bits=None, eps=1e-5, mode=0, iters=1, device=0):
    L = lib()
    W = np.ascontiguousarray(W_bits)
    N, K = W.shape
    x = np.ascontiguousarray(x, dtype=np.float32)
    B = x.shape[0]
    cols = N // 2 if mode == 2 else N
    y = np.zeros((B, cols), np.float32) if y_in is None else np.ascontiguousarray(y_in, dtype=np.float32).copy()
    nw = np.ascontiguousarray(norm_w_bits) if norm_w_bits is not None else None
    ms = C.c_float(0)
    if L.b2l_op_gemv(device, _p(W), _p(x), _p(y), _p(nw), eps, B, N, K, mode, iters, C.byref(ms)) != 0:
        raise B2lError("op_gemv: " + L.b2l_last_error(None).decode())
    return y, ms.value


def op_argmax(x, device=0):
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape[0], dtype=np.int32)
    if L.b2l_op_argmax(device, _p(x), x.shape[0], x.shape[1], _p(out)) != 0:
        raise B2lError("op_argmax: " + L.b2l_last_error(None).decode())
    return out


def op_gemm_bf16(A_bits, W_bits, epilogue=0, c_in=None, iters=0, device=0):
    """C = A[M][K] @ W[N][K]^T on the tcgen05 tensor cores. Returns (C fp32, ms per launch or 0)."""
    L = lib()
    A = np.ascontiguousarray(A_bits)
    W = np.ascontiguousarray(W_bits)
    M, K = A.shape
    N = W.shape[0]
    cols = N // 2 if epilogue == 3 else N
    Cm = np.zeros((M, cols), np.float32) if c_in is None else np.ascontiguousarray(c_in, dtype=np.float32).copy()
    ms = C.c_float(0)
    if L.b2l_op_gemm_bf16(device, _p(A), _p(W), _p(Cm), M, N, K, epilogue, iters, C.byref(ms)) != 0:
        raise B2lError("op_gemm_bf16: " + L.b2l_last_error(None).decode())
    return Cm, ms.value

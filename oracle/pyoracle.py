"""ctypes front end of oracle/liboracle.so (and oracle/_ref/liboracle_ref.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. Nothing under gabby_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

ORC_KV_BF16 = 1
ORC_ACT_BF16 = 2
ORC_QP_BF16 = 4


class OrcParams(C.Structure):
    _fields_ = [
        ("hidden_size", C.c_int32), ("intermediate_size", C.c_int32), ("num_layers", C.c_int32),
        ("num_heads", C.c_int32), ("num_kv_heads", C.c_int32), ("head_dim", C.c_int32),
        ("vocab_size", C.c_int32), ("tie_word_embeddings", C.c_int32), ("rms_norm_eps", C.c_float),
        ("rope_theta", C.c_double), ("rope_llama3", C.c_int32), ("rope_factor", C.c_double),
        ("rope_low_freq_factor", C.c_double), ("rope_high_freq_factor", C.c_double),
        ("rope_original_max_position", C.c_int32), ("max_seq_len", C.c_int32),
    ]


def params_from_arch(a, max_seq_len: int) -> OrcParams:
    return OrcParams(a.hidden_size, a.intermediate_size, a.num_hidden_layers, a.num_attention_heads,
                     a.num_key_value_heads, a.head_dim, a.vocab_size, int(a.tie_word_embeddings),
                     a.rms_norm_eps, a.rope_theta, 1, a.rope_factor, a.rope_low_freq_factor,
                     a.rope_high_freq_factor, a.rope_original_max_position, max_seq_len)


def build(ref: bool = False) -> None:
    """Compile the checker (building it is not using it)."""
    targets = ["all"] + (["ref"] if ref and os.path.isdir("/root/reference/src") else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def _bind(lib):
    vp, i32p, f32p, u16p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_uint16)
    lib.orc_model_create.restype = vp
    lib.orc_model_create.argtypes = [C.POINTER(OrcParams)]
    lib.orc_model_set_tensor.argtypes = [vp, C.c_char_p, vp, C.c_int64]
    lib.orc_model_check.argtypes = [vp]
    lib.orc_model_destroy.argtypes = [vp]
    lib.orc_seq_create.restype = vp
    lib.orc_seq_create.argtypes = [vp, C.c_int]
    lib.orc_seq_reset.argtypes = [vp]
    lib.orc_seq_set_flags.argtypes = [vp, C.c_int]
    lib.orc_seq_len.argtypes = [vp]
    lib.orc_seq_destroy.argtypes = [vp]
    lib.orc_seq_forward.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    lib.orc_argmax.restype = C.c_int32
    lib.orc_argmax.argtypes = [vp, C.c_int64]
    lib.orc_greedy.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    lib.orc_rope_table.argtypes = [C.POINTER(OrcParams), C.c_int, vp]
    lib.orc_rope_inv_freq.argtypes = [C.POINTER(OrcParams), vp]
    lib.orc_synth_tensor.argtypes = [C.c_uint32, C.c_int64, C.c_float, C.c_float, vp]
    lib.orc_num_threads.restype = C.c_int
    lib.orc_set_num_threads.argtypes = [C.c_int]
    return lib


_LIB = None
_REF = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = _bind(C.CDLL(path))
    return _LIB


def ref_lib():
    """oracle/_ref/liboracle_ref.so: same forward, loading through the reference's parsers."""
    global _REF
    if _REF is None:
        path = os.path.join(HERE, "_ref", "liboracle_ref.so")
        if not os.path.exists(path):
            build(ref=True)
        if not os.path.exists(path):
            return None
        r = _bind(C.CDLL(path))
        r.orc_ref_load.restype = C.c_void_p
        r.orc_ref_load.argtypes = [C.c_char_p, C.c_int]
        r.orc_ref_error.restype = C.c_char_p
        r.orc_ref_error.argtypes = [C.c_void_p]
        r.orc_ref_model.restype = C.c_void_p
        r.orc_ref_model.argtypes = [C.c_void_p]
        r.orc_ref_stub_generate.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        r.orc_ref_free.argtypes = [C.c_void_p]
        _REF = r
    return _REF


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleModel:
    """Owns an orc_model; `tensors` maps hf_name -> contiguous uint16 array of bf16 bits."""

    def __init__(self, arch, tensors: dict, max_seq_len: int, _lib=None, _handle=None, _owner=None):
        self.L = _lib or lib()
        self.arch = arch
        self.params = params_from_arch(arch, max_seq_len)
        self._keep = []
        self._owner = _owner
        if _handle is not None:
            self.h = _handle
            return
        self.h = self.L.orc_model_create(C.byref(self.params))
        if not self.h:
            raise ValueError("oracle: unsupported shapes")
        for name, arr in tensors.items():
            a = np.ascontiguousarray(arr).reshape(-1)
            assert a.dtype == np.uint16
            self._keep.append(a)
            if self.L.orc_model_set_tensor(self.h, name.encode(), _ptr(a), a.size) != 0:
                raise ValueError(f"oracle: bad tensor {name} ({a.size})")
        if self.L.orc_model_check(self.h) != 0:
            raise ValueError("oracle: missing tensors")

    @classmethod
    def from_dir(cls, model_dir: str, arch, max_seq_len: int):
        from gabby_b200 import synth  # synth is a data generator, not the product path
        t = synth.read_safetensors(os.path.join(model_dir, "model.safetensors"))
        return cls(arch, {k: v[1] for k, v in t.items()}, max_seq_len)

    @classmethod
    def from_dir_via_reference(cls, model_dir: str, arch, max_seq_len: int):
        """Load through gabby's own LoadConfig/Safetensors/json parsers (oracle/_ref)."""
        r = ref_lib()
        if r is None:
            raise RuntimeError("oracle/_ref/liboracle_ref.so not built (needs /root/reference)")
        h = r.orc_ref_load(model_dir.encode(), max_seq_len)
        err = r.orc_ref_error(h)
        if err:
            msg = err.decode()
            r.orc_ref_free(h)
            raise RuntimeError("reference loader: " + msg)
        return cls(arch, {}, max_seq_len, _lib=r, _handle=r.orc_ref_model(h), _owner=(r, h))

    def seq(self, flags: int = 0) -> "OracleSeq":
        return OracleSeq(self, flags)

    def close(self):
        if self._owner is not None:
            r, h = self._owner
            r.orc_ref_free(h)
            self._owner = None
        elif self.h:
            self.L.orc_model_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OracleSeq:
    def __init__(self, model: OracleModel, flags: int):
        self.m = model
        self.L = model.L
        self.h = self.L.orc_seq_create(model.h, flags)

    def __len__(self):
        return self.L.orc_seq_len(self.h)

    def reset(self):
        self.L.orc_seq_reset(self.h)

    def set_flags(self, flags: int):
        self.L.orc_seq_set_flags(self.h, flags)

    def forward(self, tokens, logits_all: bool = False, want_hidden: bool = False):
        """-> (logits [rows, V] fp32, hidden [(L+2), n, H] fp32 or None)"""
        tok = np.ascontiguousarray(tokens, dtype=np.int32)
        n = tok.size
        a = self.m.arch
        logits = np.empty((n if logits_all else 1, a.vocab_size), dtype=np.float32)
        hidden = np.empty((a.num_hidden_layers + 2, n, a.hidden_size), dtype=np.float32) if want_hidden else None
        rc = self.L.orc_seq_forward(self.h, _ptr(tok), n, int(logits_all), _ptr(logits),
                                    _ptr(hidden) if want_hidden else None)
        if rc != 0:
            raise RuntimeError("oracle forward failed (capacity or bad token)")
        return logits, hidden

    def greedy(self, prompt, n_new: int):
        """-> (ids[n_new] int32, margins[n_new] fp32)"""
        tok = np.ascontiguousarray(prompt, dtype=np.int32)
        ids = np.empty(n_new, dtype=np.int32)
        margins = np.empty(n_new, dtype=np.float32)
        if self.L.orc_greedy(self.h, _ptr(tok), tok.size, n_new, _ptr(ids), _ptr(margins)) != 0:
            raise RuntimeError("oracle greedy failed")
        return ids, margins

    def close(self):
        if self.h:
            self.L.orc_seq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rope_table(arch, max_pos: int) -> np.ndarray:
    p = params_from_arch(arch, max_pos)
    out = np.empty((max_pos, arch.head_dim // 2, 2), dtype=np.float32)
    lib().orc_rope_table(C.byref(p), max_pos, _ptr(out))
    return out


def rope_inv_freq(arch) -> np.ndarray:
    p = params_from_arch(arch, 1)
    out = np.empty(arch.head_dim // 2, dtype=np.float32)
    lib().orc_rope_inv_freq(C.byref(p), _ptr(out))
    return out


def synth_tensor(tensor_seed: int, n: int, scale: float, offset: float) -> np.ndarray:
    out = np.empty(n, dtype=np.uint16)
    lib().orc_synth_tensor(tensor_seed, n, scale, offset, _ptr(out))
    return out

"""ctypes bindings of include/gabby_b200_host.h (the C view of the C++ host layer)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libgabby_host.so")

SYMBOLS = [
    "gb_last_error", "gb_rope_table", "gb_generator_load", "gb_generator_free", "gb_generator_generate",
    "gb_generator_generate_detailed", "gb_generator_generate_ids", "gb_generator_sched_stats", "gb_generator_engine", "gb_params_from_dir", "gb_params_from_json",
    "gb_checkpoint_info", "gb_checkpoint_tensor", "gb_kv_create", "gb_kv_free", "gb_kv_new_sequence", "gb_kv_reserve",
    "gb_kv_release", "gb_kv_table", "gb_kv_free_pages", "gb_tokenizer_create", "gb_tokenizer_free", "gb_tokenize",
    "gb_detokenize", "gb_chat_prompt", "gb_argmax",
    "gb_sched_create_b2l", "gb_sched_create_fake", "gb_sched_free", "gb_sched_submit", "gb_sched_step", "gb_sched_drain",
    "gb_sched_result", "gb_sched_stats",
]


class GbParams(C.Structure):
    _fields_ = [
        ("hidden_size", C.c_int32), ("intermediate_size", C.c_int32), ("num_hidden_layers", C.c_int32),
        ("num_attention_heads", C.c_int32), ("num_key_value_heads", C.c_int32), ("head_dim", C.c_int32),
        ("vocab_size", C.c_int32), ("tie_word_embeddings", C.c_int32), ("max_position_embeddings", C.c_int32),
        ("bos_token_id", C.c_int32), ("n_eos", C.c_int32), ("eos_token_ids", C.c_int32 * 8),
        ("rope_llama3", C.c_int32), ("rope_original_max_position", C.c_int32), ("rms_norm_eps", C.c_float),
        ("rope_theta", C.c_double), ("rope_factor", C.c_double), ("rope_low_freq_factor", C.c_double),
        ("rope_high_freq_factor", C.c_double),
    ]


_LIB = None


class HostError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise HostError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)  # links libb2l.so through its $ORIGIN rpath
        vp, ip = C.c_void_p, C.POINTER(C.c_int)
        L.gb_last_error.restype = C.c_char_p
        L.gb_rope_table.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, vp]
        L.gb_generator_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.gb_generator_free.argtypes = [vp]
        L.gb_generator_free.restype = None
        L.gb_generator_generate.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        L.gb_generator_generate_detailed.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, ip, ip, ip]
        L.gb_generator_generate_ids.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, ip, ip]
        L.gb_generator_sched_stats.argtypes = [vp, vp]
        L.gb_generator_engine.argtypes = [vp]
        L.gb_generator_engine.restype = vp
        L.gb_params_from_dir.argtypes = [C.c_char_p, C.POINTER(GbParams)]
        L.gb_params_from_json.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(GbParams)]
        L.gb_checkpoint_info.argtypes = [C.c_char_p, ip, ip]
        L.gb_checkpoint_tensor.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int64), ip, C.c_char_p,
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.gb_kv_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.gb_kv_free.argtypes = [vp]
        L.gb_kv_free.restype = None
        L.gb_kv_new_sequence.argtypes = [vp, ip]
        L.gb_kv_reserve.argtypes = [vp, C.c_int, C.c_int]
        L.gb_kv_release.argtypes = [vp, C.c_int]
        L.gb_kv_table.argtypes = [vp, C.c_int, vp, C.c_int, ip]
        L.gb_kv_free_pages.argtypes = [vp]
        L.gb_tokenizer_create.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.gb_tokenizer_free.argtypes = [vp]
        L.gb_tokenizer_free.restype = None
        L.gb_tokenize.argtypes = [vp, C.c_char_p, vp, C.c_int, ip]
        L.gb_detokenize.argtypes = [vp, vp, C.c_int, C.c_char_p, C.c_int]
        L.gb_chat_prompt.argtypes = [vp, C.c_char_p, C.c_char_p, vp, C.c_int, ip]
        L.gb_argmax.argtypes = [vp, C.c_int64]
        L.gb_argmax.restype = C.c_int32
        L.gb_sched_create_b2l.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.gb_sched_create_fake.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.gb_sched_free.argtypes = [vp]
        L.gb_sched_free.restype = None
        L.gb_sched_submit.argtypes = [vp, vp, C.c_int, C.c_int, ip]
        L.gb_sched_step.argtypes = [vp, ip]
        L.gb_sched_drain.argtypes = [vp]
        L.gb_sched_result.argtypes = [vp, C.c_int, vp, C.c_int, ip, ip, ip]
        L.gb_sched_stats.argtypes = [vp, vp]
        _LIB = L
    return _LIB


def _ck(rc):
    if rc != 0:
        raise HostError(lib().gb_last_error().decode())


def rope_table(arch, max_pos: int) -> np.ndarray:
    """(cos, sin) table from the host layer's RopeTable (gabby_b200/host/params.cc)."""
    out = np.empty((max_pos, arch.head_dim // 2, 2), dtype=np.float32)
    _ck(lib().gb_rope_table(arch.rope_theta, 1, arch.rope_factor, arch.rope_low_freq_factor, arch.rope_high_freq_factor,
                            arch.rope_original_max_position, arch.head_dim, max_pos, out.ctypes.data_as(C.c_void_p)))
    return out


def params_from_dir(model_dir: str) -> GbParams:
    p = GbParams()
    _ck(lib().gb_params_from_dir(model_dir.encode(), C.byref(p)))
    return p


def params_from_json(config_json: str, gen_json: str = "") -> GbParams:
    p = GbParams()
    _ck(lib().gb_params_from_json(config_json.encode(), gen_json.encode(), C.byref(p)))
    return p


def checkpoint_info(model_dir: str):
    n, f = C.c_int(0), C.c_int(0)
    _ck(lib().gb_checkpoint_info(model_dir.encode(), C.byref(n), C.byref(f)))
    return n.value, f.value


def checkpoint_tensor(model_dir: str, name: str):
    """-> (shape tuple, dtype str, nbytes, fnv1a64 of the bytes)"""
    shape = (C.c_int64 * 4)()
    nd = C.c_int(0)
    dt = C.create_string_buffer(8)
    nb, h = C.c_uint64(0), C.c_uint64(0)
    _ck(lib().gb_checkpoint_tensor(model_dir.encode(), name.encode(), shape, C.byref(nd), dt, C.byref(nb), C.byref(h)))
    return tuple(shape[i] for i in range(nd.value)), dt.value.decode(), nb.value, h.value


class KvAllocator:
    def __init__(self, num_pages, page_size, max_blocks):
        self.h = C.c_void_p()
        _ck(lib().gb_kv_create(num_pages, page_size, max_blocks, C.byref(self.h)))
        self.max_blocks = max_blocks

    def new_sequence(self) -> int:
        s = C.c_int(0)
        _ck(lib().gb_kv_new_sequence(self.h, C.byref(s)))
        return s.value

    def reserve(self, seq, total_tokens):
        _ck(lib().gb_kv_reserve(self.h, seq, total_tokens))

    def release(self, seq):
        _ck(lib().gb_kv_release(self.h, seq))

    def table(self, seq) -> np.ndarray:
        out = np.zeros(self.max_blocks, np.int32)
        n = C.c_int(0)
        _ck(lib().gb_kv_table(self.h, seq, out.ctypes.data_as(C.c_void_p), self.max_blocks, C.byref(n)))
        return out[: n.value]

    def free_pages(self) -> int:
        return lib().gb_kv_free_pages(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().gb_kv_free(self.h)
            self.h = None


class Tokenizer:
    def __init__(self, tokenizer_json: str = ""):
        self.h = C.c_void_p()
        _ck(lib().gb_tokenizer_create(tokenizer_json.encode(), C.byref(self.h)))

    def tokenize(self, text: str):
        out = np.zeros(max(16, 4 * len(text.encode()) + 16), np.int32)
        n = C.c_int(0)
        _ck(lib().gb_tokenize(self.h, text.encode(), out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)))
        return out[: n.value].tolist()

    def detokenize_bytes(self, ids) -> bytes:
        a = np.ascontiguousarray(ids, np.int32)
        buf = C.create_string_buffer(16 * max(1, a.size) + 16)
        _ck(lib().gb_detokenize(self.h, a.ctypes.data_as(C.c_void_p), a.size, buf, len(buf)))
        return buf.value

    def detokenize(self, ids):
        return self.detokenize_bytes(ids).decode(errors="replace")

    def chat_prompt(self, system: str, user: str):
        out = np.zeros(4 * (len(system.encode()) + len(user.encode())) + 64, np.int32)
        n = C.c_int(0)
        _ck(lib().gb_chat_prompt(self.h, system.encode(), user.encode(), out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)))
        return out[: n.value].tolist()

    def __del__(self):
        if getattr(self, "h", None):
            lib().gb_tokenizer_free(self.h)
            self.h = None


class Generator:
    """gabby::inference::Llama3Generator behind its C view."""

    def __init__(self, model_dir: str, device=0, max_positions=2048, max_new_tokens=256):
        self.h = C.c_void_p()
        _ck(lib().gb_generator_load(model_dir.encode(), device, max_positions, max_new_tokens, C.byref(self.h)))

    def generate(self, system: str, user: str) -> str:
        buf = C.create_string_buffer(1 << 16)
        _ck(lib().gb_generator_generate(self.h, system.encode(), user.encode(), buf, len(buf)))
        return buf.value.decode(errors="replace")

    def generate_ids(self, prompt, max_new_tokens: int, device_loop: bool = False):
        p = np.ascontiguousarray(prompt, np.int32)
        out = np.zeros(max_new_tokens + 1, np.int32)
        n, fin = C.c_int(0), C.c_int(0)
        _ck(lib().gb_generator_generate_ids(self.h, p.ctypes.data_as(C.c_void_p), p.size, max_new_tokens, int(device_loop),
                                            out.ctypes.data_as(C.c_void_p), out.size, C.byref(n), C.byref(fin)))
        return out[: min(n.value, out.size)].copy(), {1: "stop", 2: "length"}.get(fin.value, "none")

    def generate_detailed(self, system: str, user: str, max_tokens: int = 0, cap: int = 1 << 16):
        """-> (text, prompt_tokens, completion_tokens, finish_reason): what an OpenAI-style response reports."""
        buf = C.create_string_buffer(cap)
        pt, ct, fin = C.c_int(0), C.c_int(0), C.c_int(0)
        _ck(lib().gb_generator_generate_detailed(self.h, system.encode(), user.encode(), max_tokens, buf, cap, C.byref(pt), C.byref(ct), C.byref(fin)))
        return buf.value.decode(errors="replace"), pt.value, ct.value, {1: "stop", 2: "length"}.get(fin.value, "none")

    def sched_stats(self):
        out = np.zeros(8, np.int64)
        _ck(lib().gb_generator_sched_stats(self.h, out.ctypes.data_as(C.c_void_p)))
        return dict(zip(["steps", "prefill_calls", "decode_calls", "prefill_tokens", "decode_tokens", "preemptions", "max_concurrent", "free_pages"], out.tolist()))

    def close(self):
        if getattr(self, "h", None):
            lib().gb_generator_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scheduler:
    """Continuous batching (gabby_b200/host/scheduler.h). `engine`: a gabby_b200._capi.Engine created with
    max_batch >= max_batch here (its page pool is managed by this scheduler), or None for the deterministic fake
    engine (next = (31 * last + 7 * position + 3) mod vocab) used by the CPU policy tests."""

    def __init__(self, engine=None, *, eos_ids=(), max_batch=8, max_positions=2048, max_prefill_tokens=2048, num_pages=0,
                 page_size=16, fake_vocab=0):
        self.L = lib()
        self.h = C.c_void_p()
        self._engine = engine   # keep the engine alive
        if engine is None:
            eos = int(eos_ids[0]) if len(eos_ids) else -1
            _ck(self.L.gb_sched_create_fake(fake_vocab, eos, max_batch, max_positions, max_prefill_tokens, num_pages, page_size,
                                            C.byref(self.h)))
        else:
            e = np.ascontiguousarray(eos_ids, dtype=np.int32)
            _ck(self.L.gb_sched_create_b2l(engine.h, e.ctypes.data_as(C.c_void_p), e.size, max_batch, max_positions,
                                           max_prefill_tokens, num_pages, page_size, C.byref(self.h)))

    def submit(self, prompt, max_new_tokens: int) -> int:
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        rid = C.c_int(-1)
        _ck(self.L.gb_sched_submit(self.h, p.ctypes.data_as(C.c_void_p), p.size, max_new_tokens, C.byref(rid)))
        return rid.value

    def step(self) -> int:
        n = C.c_int(0)
        _ck(self.L.gb_sched_step(self.h, C.byref(n)))
        return n.value

    def drain(self):
        _ck(self.L.gb_sched_drain(self.h))

    def result(self, rid: int):
        """-> (tokens, finish ('none' | 'stop' | 'length'), done)"""
        n, fin, done = C.c_int(0), C.c_int(0), C.c_int(0)
        _ck(self.L.gb_sched_result(self.h, rid, None, 0, C.byref(n), C.byref(fin), C.byref(done)))
        out = np.zeros(max(1, n.value), dtype=np.int32)
        _ck(self.L.gb_sched_result(self.h, rid, out.ctypes.data_as(C.c_void_p), out.size, C.byref(n), C.byref(fin), C.byref(done)))
        return out[:n.value].tolist(), ("none", "stop", "length")[fin.value], bool(done.value)

    def stats(self) -> dict:
        out = np.zeros(8, dtype=np.int64)
        _ck(self.L.gb_sched_stats(self.h, out.ctypes.data_as(C.c_void_p)))
        keys = ("steps", "prefill_calls", "decode_calls", "prefill_tokens", "decode_tokens", "preemptions", "max_concurrent", "free_pages")
        return dict(zip(keys, out.tolist()))

    def close(self):
        if getattr(self, "h", None):
            self.L.gb_sched_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def argmax(logits) -> int:
    a = np.ascontiguousarray(logits, np.float32)
    return int(lib().gb_argmax(a.ctypes.data_as(C.c_void_p), a.size))

"""Synthetic (random-init) Llama-3 model directories in the HF layout gabby loads.

The reference's `inference::LoadConfig` (/root/reference/src/inference/config.cc:11-28)
reads five JSON files plus a single `model.safetensors` from a snapshot directory.
This module writes such a directory with deterministic weights so that the CPU oracle,
the CUDA engine (which can also generate the same weights on-device) and HF
transformers all see the same bytes.

Weight generator ("b2l counter hash", restated bit-exactly in oracle/llama_oracle.cc and
gabby_b200/csrc/synth.cuh):

    seed_t   = fnv1a32(tensor_name) ^ (global_seed * 0x9E3779B9)
    x        = lowbias32(i + seed_t)                 # i = flat element index, uint32 wrap
    u        = float32(x >> 8) * 2^-23 - 1.0         # exact in fp32, in [-1, 1)
    w[i]     = bf16_rne(float32(offset) + u * float32(scale))

`scale`/`offset` per tensor kind: linear weights use sqrt(3/fan_in) (unit gain),
embeddings `embed_scale`, norm weights 1 + 0.1*u.
"""
from __future__ import annotations

import json
import os
import struct
from dataclasses import dataclass, asdict, replace

import numpy as np

# ----------------------------------------------------------------------------------
# architecture presets (public Llama-3.x config.json values)
# ----------------------------------------------------------------------------------


@dataclass(frozen=True)
class LlamaArch:
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    head_dim: int
    vocab_size: int
    tie_word_embeddings: bool
    rope_theta: float = 500000.0
    rope_factor: float = 32.0
    rope_low_freq_factor: float = 1.0
    rope_high_freq_factor: float = 4.0
    rope_original_max_position: int = 8192
    max_position_embeddings: int = 131072
    rms_norm_eps: float = 1e-5
    bos_token_id: int = 128000
    eos_token_ids: tuple = (128001, 128008, 128009)


PRESETS = {
    # tiny: CPU-test sized, exercises GQA (8 q / 2 kv), llama3 rope scaling, tied head
    "tiny": LlamaArch(256, 512, 2, 8, 2, 32, 1024, True, rope_original_max_position=64,
                      max_position_embeddings=2048, bos_token_id=1, eos_token_ids=(2,)),
    # tiny-untied: head_dim 128 / untied lm_head like 8B/70B
    "tiny128": LlamaArch(512, 1024, 2, 4, 2, 128, 1536, False, rope_factor=8.0,
                         rope_original_max_position=64, max_position_embeddings=2048,
                         bos_token_id=1, eos_token_ids=(2,)),
    "1b": LlamaArch(2048, 8192, 16, 32, 8, 64, 128256, True),
    "3b": LlamaArch(3072, 8192, 28, 24, 8, 128, 128256, True),
    "8b": LlamaArch(4096, 14336, 32, 32, 8, 128, 128256, False, rope_factor=8.0),
    "70b": LlamaArch(8192, 28672, 80, 64, 8, 128, 128256, False, rope_factor=8.0),
    # the per-rank shapes of Llama-3.1-8B at TP=8 as a single-GPU model (4 query heads on 1 kv head, I/8, V/8): a one-GPU proxy
    # for tuning the phases of a tensor-parallel rank (tools/tp_proxy_bench.py); not a real model
    "8b_tp8rank": LlamaArch(4096, 1792, 32, 4, 1, 128, 16032, False, rope_factor=8.0),
}


def preset(name: str, layers: int | None = None) -> LlamaArch:
    a = PRESETS[name]
    return replace(a, num_hidden_layers=layers) if layers is not None else a


# ----------------------------------------------------------------------------------
# counter-hash RNG
# ----------------------------------------------------------------------------------


def fnv1a32(s: str) -> int:
    h = 0x811C9DC5
    for b in s.encode("utf-8"):
        h = ((h ^ b) * 0x01000193) & 0xFFFFFFFF
    return h


def tensor_seed(name: str, seed: int) -> int:
    return (fnv1a32(name) ^ ((seed * 0x9E3779B9) & 0xFFFFFFFF)) & 0xFFFFFFFF


def _lowbias32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 bit pattern (uint16). No NaN inputs here."""
    b = a.astype(np.float32).view(np.uint32)
    rounding = np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))
    return ((b + rounding) >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


def gen_tensor_bits(name: str, n: int, scale: float, offset: float, seed: int,
                    chunk: int = 1 << 24) -> np.ndarray:
    """bf16 bit patterns (uint16) of tensor `name` with `n` elements."""
    out = np.empty(n, dtype=np.uint16)
    ts = np.uint32(tensor_seed(name, seed))
    sc = np.float32(scale)
    off = np.float32(offset)
    with np.errstate(over="ignore"):
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            idx = np.arange(lo, hi, dtype=np.uint64).astype(np.uint32) + ts
            x = _lowbias32(idx)
            u = (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -23) - np.float32(1.0)
            out[lo:hi] = f32_to_bf16_bits(off + u * sc)
    return out


# ----------------------------------------------------------------------------------
# tensor inventory
# ----------------------------------------------------------------------------------


def tensor_specs(a: LlamaArch, embed_scale: float = 0.05):
    """[(hf_name, shape, scale, offset)] in the order HF writes them."""
    H, I, V = a.hidden_size, a.intermediate_size, a.vocab_size
    qd = a.num_attention_heads * a.head_dim
    kvd = a.num_key_value_heads * a.head_dim

    def lin(fan_in):
        return float(np.sqrt(3.0 / fan_in))

    specs = [("model.embed_tokens.weight", (V, H), embed_scale, 0.0)]
    for l in range(a.num_hidden_layers):
        p = f"model.layers.{l}."
        specs += [
            (p + "input_layernorm.weight", (H,), 0.1, 1.0),
            (p + "self_attn.q_proj.weight", (qd, H), lin(H), 0.0),
            (p + "self_attn.k_proj.weight", (kvd, H), lin(H), 0.0),
            (p + "self_attn.v_proj.weight", (kvd, H), lin(H), 0.0),
            (p + "self_attn.o_proj.weight", (H, qd), lin(qd), 0.0),
            (p + "post_attention_layernorm.weight", (H,), 0.1, 1.0),
            (p + "mlp.gate_proj.weight", (I, H), lin(H), 0.0),
            (p + "mlp.up_proj.weight", (I, H), lin(H), 0.0),
            (p + "mlp.down_proj.weight", (H, I), lin(I), 0.0),
        ]
    specs.append(("model.norm.weight", (H,), 0.1, 1.0))
    if not a.tie_word_embeddings:
        specs.append(("lm_head.weight", (V, H), embed_scale, 0.0))
    return specs


def hf_config(a: LlamaArch) -> dict:
    return {
        "architectures": ["LlamaForCausalLM"],
        "attention_bias": False,
        "attention_dropout": 0.0,
        "bos_token_id": a.bos_token_id,
        "eos_token_id": list(a.eos_token_ids),
        "head_dim": a.head_dim,
        "hidden_act": "silu",
        "hidden_size": a.hidden_size,
        "initializer_range": 0.02,
        "intermediate_size": a.intermediate_size,
        "max_position_embeddings": a.max_position_embeddings,
        "mlp_bias": False,
        "model_type": "llama",
        "num_attention_heads": a.num_attention_heads,
        "num_hidden_layers": a.num_hidden_layers,
        "num_key_value_heads": a.num_key_value_heads,
        "pretraining_tp": 1,
        "rms_norm_eps": a.rms_norm_eps,
        "rope_scaling": {
            "factor": a.rope_factor,
            "high_freq_factor": a.rope_high_freq_factor,
            "low_freq_factor": a.rope_low_freq_factor,
            "original_max_position_embeddings": a.rope_original_max_position,
            "rope_type": "llama3",
        },
        "rope_theta": a.rope_theta,
        "tie_word_embeddings": a.tie_word_embeddings,
        "torch_dtype": "bfloat16",
        "use_cache": True,
        "vocab_size": a.vocab_size,
    }


def write_model_dir(path: str, arch: LlamaArch, seed: int = 1234, embed_scale: float = 0.05,
                    shards: int = 1) -> str:
    """Write an HF-layout snapshot dir. `shards`>1 writes model-0000i-of-0000n.safetensors
    plus model.safetensors.index.json (the layout real 3B/8B/70B checkpoints use)."""
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(hf_config(arch), f, indent=2)
    with open(os.path.join(path, "generation_config.json"), "w") as f:
        json.dump({"bos_token_id": arch.bos_token_id, "do_sample": False,
                   "eos_token_id": list(arch.eos_token_ids)}, f)
    with open(os.path.join(path, "special_tokens_map.json"), "w") as f:
        json.dump({"bos_token": "<|begin_of_text|>", "eos_token": "<|eot_id|>"}, f)
    with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
        json.dump({"bos_token": "<|begin_of_text|>", "eos_token": "<|eot_id|>",
                   "model_max_length": arch.max_position_embeddings}, f)
    with open(os.path.join(path, "tokenizer.json"), "w") as f:
        json.dump({"version": "1.0", "model": {"type": "BPE", "vocab": {}, "merges": []},
                   "added_tokens": []}, f)
    with open(os.path.join(path, "b2l_synth.json"), "w") as f:
        json.dump({"seed": seed, "embed_scale": embed_scale, "arch": asdict(arch)}, f)

    specs = tensor_specs(arch, embed_scale)
    groups = [specs] if shards <= 1 else [list(g) for g in np.array_split(np.array(specs, dtype=object), shards)]
    weight_map = {}
    for gi, group in enumerate(groups):
        fname = ("model.safetensors" if shards <= 1
                 else f"model-{gi + 1:05d}-of-{len(groups):05d}.safetensors")
        header, off = {}, 0
        for name, shape, _, _ in group:
            n = int(np.prod(shape)) * 2
            header[name] = {"dtype": "BF16", "shape": list(shape), "data_offsets": [off, off + n]}
            off += n
            weight_map[name] = fname
        header["__metadata__"] = {"format": "pt"}
        hbytes = json.dumps(header, separators=(",", ":")).encode()
        hbytes += b" " * ((8 - len(hbytes) % 8) % 8)
        with open(os.path.join(path, fname), "wb") as f:
            f.write(struct.pack("<Q", len(hbytes)))
            f.write(hbytes)
            for name, shape, scale, offset in group:
                gen_tensor_bits(name, int(np.prod(shape)), scale, offset, seed).tofile(f)
    if shards > 1:
        with open(os.path.join(path, "model.safetensors.index.json"), "w") as f:
            json.dump({"metadata": {}, "weight_map": weight_map}, f)
    return path


def read_safetensors(path: str) -> dict:
    """name -> (shape, uint16 memmap of bf16 bits). Test/bench helper."""
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen))
    mm = np.memmap(path, dtype=np.uint8, mode="r", offset=8 + hlen)
    out = {}
    for name, meta in header.items():
        if name == "__metadata__":
            continue
        assert meta["dtype"] == "BF16", meta
        b, e = meta["data_offsets"]
        out[name] = (tuple(meta["shape"]), mm[b:e].view(np.uint16))
    return out


def synth_prompt(n: int, vocab: int, bos: int, seed: int) -> np.ndarray:
    """BOS + (n-1) ids uniform in [0, min(vocab, 128000)) from the same counter hash."""
    hi = min(vocab, 128000)
    idx = np.arange(n, dtype=np.uint32) + np.uint32(tensor_seed("prompt", seed))
    with np.errstate(over="ignore"):
        ids = (_lowbias32(idx) % np.uint32(hi)).astype(np.int32)
    ids[0] = bos
    return ids


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="tiny")
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--shards", type=int, default=1)
    ap.add_argument("out")
    args = ap.parse_args()
    print(write_model_dir(args.out, preset(args.preset, args.layers), args.seed, shards=args.shards))

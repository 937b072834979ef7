"""Writes tests/golden/bench_parity.json: greedy ids of the CPU oracle for the two-layer, same-width variants of the
models bench.py times ("L=2 same-width variant"). bench.py runs the same prompts through the
CUDA path right after its timed loops and reports `parity_check` in its JSON line -- that is what lets a box on
which pytest's multi-GPU cases are skipped still witness tensor-parallel parity.

    python tests/golden/make_bench_parity.py            (CPU only; a few minutes)

Legs (weights: gabby_b200/synth.py counter hash, seed = bench.SEED):
  1b_l2   Llama-3.2-1B width, 512-token prompt (bench's own) through the bf16-activation prefill (tcgen05 GEMMs + flash
          attention on the GPU; oracle flags ORC_ACT_BF16 | ORC_QP_BF16), then 32 greedy tokens with fp32 activations
  3b_l2   Llama-3.2-3B width, 8 prompts of 256 tokens, same prefill, 8 greedy tokens per sequence
  8b_l2   Llama-3.1-8B width (the tensor-parallel leg), 2 prompts of 48 tokens through the exact-activation prefill,
          12 greedy tokens; every TP degree must reproduce these ids
Margins (oracle top-1 minus top-2 logit at every step) are stored so that a flip at a near-tie can be told from an error.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gabby_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

SEED = 20261018


def model(preset):
    arch = synth.preset(preset, 2)
    tensors = {n: po.synth_tensor(synth.tensor_seed(n, SEED), int(np.prod(s)), sc, off) for n, s, sc, off in synth.tensor_specs(arch)}
    return arch, tensors


def greedy(om, prompt, n_new, prefill_flags):
    s = om.seq(prefill_flags)
    lg, _ = s.forward(prompt)
    s.set_flags(po.ORC_KV_BF16)
    ids, margins = [], []
    for _ in range(n_new):
        top2 = np.partition(lg[0], -2)[-2:]
        margins.append(float(top2[1] - top2[0]))
        tok = int(np.argmax(lg[0]))
        ids.append(tok)
        lg, _ = s.forward([tok])
    return ids, margins


def main():
    out = {"seed": SEED, "generator": "tests/golden/make_bench_parity.py", "legs": {}}
    bf16 = po.ORC_KV_BF16 | po.ORC_ACT_BF16 | po.ORC_QP_BF16

    arch, tensors = model("1b")
    om = po.OracleModel(arch, tensors, 600)
    prompt = synth.synth_prompt(512, arch.vocab_size, arch.bos_token_id, SEED + 1)
    ids, margins = greedy(om, prompt, 33, bf16)
    out["legs"]["1b_l2"] = {"preset": "1b", "layers": 2, "prompt_lens": [512], "prompt_seeds": [SEED + 1], "prefill": "bf16-activations",
                            "ids": [ids], "margins": [margins]}
    print("1b_l2", ids[:8], min(margins), flush=True)
    om.close()

    arch, tensors = model("3b")
    om = po.OracleModel(arch, tensors, 300)
    rows_i, rows_m, seeds = [], [], []
    for i in range(8):
        prompt = synth.synth_prompt(256, arch.vocab_size, arch.bos_token_id, SEED + 10 + i)
        ids, margins = greedy(om, prompt, 9, bf16)
        rows_i.append(ids); rows_m.append(margins); seeds.append(SEED + 10 + i)
        print("3b_l2", i, ids, min(margins), flush=True)
    out["legs"]["3b_l2"] = {"preset": "3b", "layers": 2, "prompt_lens": [256] * 8, "prompt_seeds": seeds, "prefill": "bf16-activations",
                            "ids": rows_i, "margins": rows_m}
    om.close()

    arch, tensors = model("8b")
    om = po.OracleModel(arch, tensors, 100)
    rows_i, rows_m, seeds = [], [], []
    for i in range(2):
        prompt = synth.synth_prompt(48, arch.vocab_size, arch.bos_token_id, SEED + 30 + i)
        ids, margins = greedy(om, prompt, 13, po.ORC_KV_BF16)
        rows_i.append(ids); rows_m.append(margins); seeds.append(SEED + 30 + i)
        print("8b_l2", i, ids, min(margins), flush=True)
    out["legs"]["8b_l2"] = {"preset": "8b", "layers": 2, "prompt_lens": [48, 48], "prompt_seeds": seeds, "prefill": "exact",
                            "ids": rows_i, "margins": rows_m}
    om.close()

    path = os.path.join(ROOT, "tests", "golden", "bench_parity.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()

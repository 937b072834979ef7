"""Shared test helpers (CPU side): synthetic weights + oracle models from golden fixtures."""
import functools
import os

import numpy as np

from gabby_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    layers = int(g["layers"])
    arch = synth.preset(str(g["preset"]), None if layers < 0 else layers)
    return g, arch, int(g["seed"])


@functools.lru_cache(maxsize=4)
def synth_tensors(preset, layers, seed):
    arch = synth.preset(preset, layers)
    return arch, {n: synth.gen_tensor_bits(n, int(np.prod(s)), sc, off, seed)
                  for n, s, sc, off in synth.tensor_specs(arch)}


def cosine(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))


# ---- GPU-side helpers (only used by -m gpu tests, smoke and bench) -------------------------

def make_engine(arch, tensors, *, max_batch=1, max_positions=256, page_size=16, num_pages=None,
                max_prefill_tokens=256, synth_seed=None):
    """Engine with `tensors` uploaded (or generated on-device when synth_seed is given)."""
    from gabby_b200 import _capi
    from oracle import pyoracle as po
    eng = _capi.Engine(arch, po.rope_table(arch, max_positions), max_batch=max_batch, max_positions=max_positions,
                       page_size=page_size, num_pages=num_pages, max_prefill_tokens=max_prefill_tokens)
    for name, shape, scale, off in synth.tensor_specs(arch):
        if synth_seed is not None:
            eng.synth(name, shape, synth.tensor_seed(name, synth_seed), scale, off)
        else:
            eng.upload(name, tensors[name], shape)
    eng.finalize()
    return eng


def contiguous_tables(n_seq, max_blocks):
    return np.arange(n_seq * max_blocks, dtype=np.int32).reshape(n_seq, max_blocks)

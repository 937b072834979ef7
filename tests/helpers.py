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

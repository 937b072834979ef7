"""Tiny repro driver: 1B-width 2-layer model, a few megakernel tokens."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gabby_b200 import synth, _capi, _host
arch = synth.preset("1b", 2)
eng = _capi.Engine(arch, _host.rope_table(arch, 256), max_positions=256, max_prefill_tokens=64)
for name, shape, scale, off in synth.tensor_specs(arch):
    eng.synth(name, shape, synth.tensor_seed(name, 5), scale, off)
eng.finalize()
bt = np.arange(eng.max_blocks, dtype=np.int32)[None, :]
prompt = synth.synth_prompt(12, arch.vocab_size, arch.bos_token_id, 8)
first = eng.prefill([prompt], [0], bt)
ids, ms = eng.decode_loop(first, [12], bt, 3)
print("ok", ids[:, 0].tolist(), ms)

"""Short megakernel run for ncu: 1B shapes, context 512, a few tokens per launch."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gabby_b200 import synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
arch = synth.preset("1b")
eng = bench.build_engine(arch, 0, 1024)
bt = np.arange(eng.max_blocks, dtype=np.int32)[None, :]
prompt = synth.synth_prompt(512, arch.vocab_size, arch.bos_token_id, 7)
first = eng.prefill([prompt], [0], bt)
for _ in range(3):
    ids, ms = eng.decode_loop(first, [512], bt, steps)
print("ms/token", ms / steps, ids[:4, 0].tolist())

"""Debug: dump the raw per-CTA megakernel timeline (b2l_debug_mega_profile) of one token to an .npz for offline analysis.
usage: python tools/mega_prof_dump.py OUT.npz [context]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gabby_b200 import synth

arch = synth.preset("1b")
eng = bench.build_engine(arch, 0, 1024)
bt = np.arange(eng.max_blocks, dtype=np.int32)[None, :]
CTX = int(sys.argv[2]) if len(sys.argv) > 2 else 512
prompt = synth.synth_prompt(CTX, arch.vocab_size, arch.bos_token_id, 7)
first = eng.prefill([prompt], [0], bt)
eng.mega_profile(True)
eng.decode_loop(first, [CTX], bt, 8)
runs = []
for _ in range(3):
    ids, ms = eng.decode_loop(first, [CTX], bt, 64)
    ns, types = eng.mega_profile(True)
    runs.append(ns.copy())
np.savez_compressed(sys.argv[1], ns=np.stack(runs), types=types, ms_per_token=ms / 64)
print("ms/token", ms / 64)

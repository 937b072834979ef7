"""Debug: per-phase timeline of the decode megakernel (Llama-3.2-1B shapes, context 512)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gabby_b200 import synth

arch = synth.preset("1b")
eng = bench.build_engine(arch, 0, 1024)
bt = np.arange(eng.max_blocks, dtype=np.int32)[None, :]
prompt = synth.synth_prompt(512, arch.vocab_size, arch.bos_token_id, 7)
first = eng.prefill([prompt], [0], bt)
eng.mega_profile(True)
eng.decode_loop(first, [512], bt, 8)
ids, ms = eng.decode_loop(first, [512], bt, 64)
ns, types = eng.mega_profile(True)
n = len(types)
ns = ns[:, :n].astype(np.int64)
print("ms/token", ms / 64)
names = ["qkv", "attn", "o", "gateup", "down", "lmhead"]
print("token span us (cta0, first phase entry -> last phase end)", (ns[3, -1] - ns[0, 0]) / 1e3)
for k in range(6):
    sel = types == k
    if not sel.any():
        continue
    def avg(a):
        return a[sel].mean() / 1e3
    print(f"{names[k]:7s} n={sel.sum():3d} | cta0: barrier {avg(ns[1]-ns[0]):6.2f}  xload {avg(ns[2]-ns[1]):6.2f}  rows {avg(ns[3]-ns[2]):7.2f}"
          f"  (weight-wait {ns[8][sel].mean()/1.965e3:6.2f}) | ctaL: barrier {avg(ns[5]-ns[4]):6.2f}  xload {avg(ns[6]-ns[5]):6.2f}  rows {avg(ns[7]-ns[6]):7.2f}"
          f" | total {((ns[3]-ns[0])[sel].sum())/1e3:8.1f} us")

sel = types == 1
print("attention item breakdown, cta0 tid0 (us): q-rope / loads+scores / sub-slot merge / smem+barrier / CTA merge+store")
print(" ".join(f"{ns[r][sel].mean()/1.965e3:6.2f}" for r in (9, 10, 11, 12, 13)))

"""Debug: per-phase timeline of the decode megakernel (1B, ctx 512)."""
import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench
from gabby_b200 import synth
arch = synth.preset("1b")
eng = bench.build_engine(arch, 0, 1024)
bt = np.arange(eng.max_blocks, dtype=np.int32)[None, :]
prompt = synth.synth_prompt(512, arch.vocab_size, arch.bos_token_id, 7)
first = eng.prefill([prompt], [0], bt)
eng.mega_profile(True)
eng.decode_loop(first, [512], bt, 8)
ids, ms = eng.decode_loop(first, [512], bt, 32)
ns, types = eng.mega_profile(True)
print("ms/token", ms / 32)
names = ["qkv", "attn", "o", "gateup", "down", "lmhead"]
t0 = ns[0, 0]
work0 = ns[0, 1:] - ns[1, :-1]      # CTA0: phase start (prev barrier exit) -> phase end
wait0 = ns[1, 1:] - ns[0, 1:]       # CTA0: barrier wait
workL = ns[2, 1:] - ns[3, :-1]
waitL = ns[3, 1:] - ns[2, 1:]
print("token total us (cta0)", (ns[1, -1] - t0) / 1e3)
for k in range(6):
    sel = types == k
    print(f"{names[k]:7s} n={sel.sum():3d}  cta0 work {work0[sel].mean()/1e3:7.2f} wait {wait0[sel].mean()/1e3:6.2f} | ctaL work {workL[sel].mean()/1e3:7.2f} wait {waitL[sel].mean()/1e3:6.2f}  sum {(work0[sel].sum()+wait0[sel].sum())/1e3:8.1f} us")
print("first 12 phases cta0 work/wait us:", [(names[types[i]], round(work0[i]/1e3,1), round(wait0[i]/1e3,1)) for i in range(12)])

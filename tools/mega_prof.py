"""Debug: per-phase timeline of the decode megakernel (Llama-3.2-1B shapes, context 512)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gabby_b200 import synth

arch = synth.preset("1b")
eng = bench.build_engine(arch, 0, 1024)
bt = np.arange(eng.max_blocks, dtype=np.int32)[None, :]
CTX = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prompt = synth.synth_prompt(CTX, arch.vocab_size, arch.bos_token_id, 7)
first = eng.prefill([prompt], [0], bt)
eng.mega_profile(True)
eng.decode_loop(first, [CTX], bt, 8)
ids, ms = eng.decode_loop(first, [CTX], bt, 64)
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

# hand-off: end of the previous phase's rows (later of the two sampled CTAs) -> input vector ready in this phase
prev_end = np.maximum(ns[3], ns[7])
for k in range(6):
    idx = np.nonzero(types == k)[0]
    idx = idx[idx > 0]
    if len(idx) == 0:
        continue
    h0 = (ns[2][idx] - prev_end[idx - 1]) / 1e3
    hL = (ns[6][idx] - prev_end[idx - 1]) / 1e3
    print(f"handoff into {names[k]:7s}: cta0 {h0.mean():6.2f} us   ctaL {hL.mean():6.2f} us   (previous phase end spread |cta0-ctaL| {np.abs(ns[3][idx-1]-ns[7][idx-1]).mean()/1e3:5.2f} us)")

# all CTAs: when did each one have its input / finish its rows, relative to the first CTA to finish the phase before
G = 148
ready = ns[16:16 + G, :n]
end = ns[176:176 + G, :n]
for k in range(6):
    idx = np.nonzero(types == k)[0]
    idx = idx[idx > 1]
    if len(idx) == 0:
        continue
    rows = []
    for i in idx:
        e_prev = end[:, i - 1][end[:, i - 1] > 0]
        if k == 1:
            e_this = end[:, i][end[:, i] > 0]
            rows.append((0, 0, 0, (e_this.max() - e_this.min()) / 1e3, (e_this.max() - e_prev.max()) / 1e3))
            continue
        r = ready[:, i][ready[:, i] > 0]
        e_this = end[:, i][end[:, i] > 0]
        rows.append(((e_prev.max() - e_prev.min()) / 1e3, (r.min() - e_prev.max()) / 1e3, (r.max() - e_prev.max()) / 1e3,
                     (e_this.max() - e_this.min()) / 1e3, (e_this.max() - e_prev.max()) / 1e3))
    m = np.array(rows).mean(axis=0)
    print(f"all CTAs, {names[k]:7s}: prev-phase end spread {m[0]:5.2f} | input ready after last prev end: first CTA {m[1]:5.2f}, last CTA {m[2]:5.2f} | this-phase end spread {m[3]:5.2f} | phase length (last end -> last end) {m[4]:6.2f} us")

for k in (0, 2, 3, 4, 5):
    sel = types == k
    if sel.any() and ns[12][sel].mean() > 0:
        print(f"{names[k]:7s} rounds {ns[12][sel].mean():4.1f}: per phase (us) stage-wait {ns[9][sel].mean()/1.965e3:6.2f}  lds+fma {ns[10][sel].mean()/1.965e3:6.2f}  butterfly+epilogue {ns[11][sel].mean()/1.965e3:6.2f}")
sel = types == 1
print("attention item breakdown, cta0 tid0 (us): q-rope / loads+scores / sub-slot merge / smem+barrier / CTA merge+store")
print(" ".join(f"{ns[r][sel].mean()/1.965e3:6.2f}" for r in (9, 10, 11, 12, 13)))


# per warp of CTA 0 (builds with -DMEGA_PROF_WARP; cycle counter of its SM, 1.965 GHz): us after the warp entered the phase
if ns.shape[0] >= 16 + 320 + 64 and ns[336:400].any():
    labels = ["input ready", "1st stage landed", "1st group done", "rows done", "barrier passed", "phase left", "before 1st wait"]
    for k in (0, 2, 3, 4, 5):
        sel = types == k
        if not sel.any():
            continue
        print(names[k])
        base = ns[336:344][:, sel]
        for i, lab in enumerate(labels):
            v = ((ns[336 + 8 * (i + 1):344 + 8 * (i + 1)][:, sel] - base).mean(axis=1)) / 1.965e3
            print(f"   {lab:18s} " + " ".join(f"{x:6.2f}" for x in v))

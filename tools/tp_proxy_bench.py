"""Debug: batch-1 megakernel decode of a preset at a long context on ONE GPU (KV = zeros), e.g. the per-rank shapes of
Llama-3.1-8B at TP=8 (`8b_tp8rank`: what one tensor-parallel rank computes, minus the NVLink exchange).
usage: python tools/tp_proxy_bench.py [preset] [context] [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gabby_b200 import _capi, _host, synth

preset = sys.argv[1] if len(sys.argv) > 1 else "8b_tp8rank"
ctx0 = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 64
arch = synth.preset(preset)
max_positions = ctx0 + 3 * K + 64
eng = _capi.Engine(arch, _host.rope_table(arch, max_positions), max_batch=1, max_positions=max_positions, page_size=bench.PAGE,
                   max_prefill_tokens=64, device=0)
for name, shape, scale, off in synth.tensor_specs(arch):
    eng.synth(name, shape, synth.tensor_seed(name, bench.SEED), scale, off)
eng.finalize()
info = eng.info()
bt = np.arange(eng.max_blocks, dtype=np.int32).reshape(1, eng.max_blocks)
tok = synth.synth_prompt(2, arch.vocab_size, arch.bos_token_id, bench.SEED + 2)[1:]
eng.decode_loop(tok, [ctx0], bt, 8)
_, ms = eng.decode_loop(tok, [ctx0], bt, K)
kv_per_tok = 2 * arch.num_hidden_layers * arch.num_key_value_heads * arch.head_dim * 2
gb = (info.stream_bytes_per_token + (ctx0 + K / 2) * kv_per_tok) / 1e9
print(f"{preset} ctx {ctx0} decode_mode {info.decode_mode}: {ms / K:.4f} ms/token, {gb / (ms / K * 1e-3):.0f} GB/s ({gb:.3f} GB/token), "
      f"{ms / K * 1e3 / (arch.num_hidden_layers * 5 + 1):.2f} us per phase")
if len(sys.argv) > 4:
    eng.mega_profile(True)
    eng.decode_loop(tok, [ctx0], bt, 8)
    ns, types = eng.mega_profile(True)
    n = len(types)
    ns = ns[:, :n].astype(np.int64)
    end = ns[176:176 + 148]
    names = ["qkv", "attn", "o", "gateup", "down", "lmhead"]
    for k in range(6):
        idx = np.nonzero(types == k)[0]
        idx = idx[idx > 1]
        if len(idx) == 0:
            continue
        ln = [(end[:, i][end[:, i] > 0].max() - end[:, i - 1][end[:, i - 1] > 0].max()) / 1e3 for i in idx]
        print(f"  {names[k]:7s} phase length {np.mean(ln):6.2f} us")

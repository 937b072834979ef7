"""Batch-B decode step time of the multi-kernel path at a short and a long context (3B shapes by default):
separates the projection (weight-stream) cost from the attention (KV-stream) cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gabby_b200 import synth, _capi
from oracle import pyoracle as po   # rope table only (tools/ is test infrastructure)

name = sys.argv[1] if len(sys.argv) > 1 else "3b"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
arch = synth.preset(name, None)
contexts = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [16, 2048]   # e.g. "2048" alone for an ncu capture
for S in contexts:
    maxpos = S + 64
    eng = _capi.Engine(arch, po.rope_table(arch, maxpos), max_batch=B, max_positions=maxpos, page_size=16, max_prefill_tokens=B * S)
    for n, shape, scale, off in synth.tensor_specs(arch):
        eng.synth(n, shape, synth.tensor_seed(n, 1), scale, off)
    eng.finalize()
    bt = np.arange(B * eng.max_blocks, dtype=np.int32).reshape(B, eng.max_blocks)
    prompts = [synth.synth_prompt(S, arch.vocab_size, arch.bos_token_id, i) for i in range(B)]
    first = eng.prefill(prompts, [0] * B, bt)
    eng.decode_loop(first, [S] * B, bt, 3)
    _, ms = eng.decode_loop(first, [S] * B, bt, 20)
    info = eng.info()
    print(f"{name} B={B} ctx={S}: {ms / 20:.3f} ms/step  weights {info.stream_bytes_per_token / 1e9:.2f} GB -> {info.stream_bytes_per_token / (ms / 20) / 1e9:.0f} GB/s", flush=True)
    eng.close()

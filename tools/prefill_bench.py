"""Prefill throughput: tcgen05 GEMM path vs the chunked decode-kernel path (1B shapes)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gabby_b200 import synth, _capi, _host
arch = synth.preset(sys.argv[1] if len(sys.argv) > 1 else "1b")
B, S = int(sys.argv[2]) if len(sys.argv) > 2 else 1, int(sys.argv[3]) if len(sys.argv) > 3 else 2048
eng = _capi.Engine(arch, _host.rope_table(arch, S + 64), max_batch=B, max_positions=S + 64, page_size=16, max_prefill_tokens=B * S)
for name, shape, scale, off in synth.tensor_specs(arch):
    eng.synth(name, shape, synth.tensor_seed(name, 1), scale, off)
eng.finalize()
bt = np.arange(B * eng.max_blocks, dtype=np.int32).reshape(B, eng.max_blocks)
prompts = [synth.synth_prompt(S, arch.vocab_size, arch.bos_token_id, 10 + i) for i in range(B)]
for mode in (1, 0) if S * B <= 4096 else (1,):
    eng.set_prefill_mode(mode)
    eng.prefill(prompts, [0] * B, bt)
    t0 = time.perf_counter(); first = eng.prefill(prompts, [0] * B, bt); dt = time.perf_counter() - t0
    print(f"prefill mode {mode}: {B} x {S} tokens in {dt*1e3:.1f} ms = {B*S/dt:,.0f} tok/s   first ids {first.tolist()[:4]}")

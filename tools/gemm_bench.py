"""tcgen05 GEMM throughput on the Llama-3.2-3B prefill shapes (BASELINE configs[2]: M = 8 x 2048 tokens)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gabby_b200 import _capi
rng = np.random.default_rng(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
peak = json.load(open("MEASURED_PEAKS.json"))["bf16_tflops"] if os.path.exists("MEASURED_PEAKS.json") else 1590.0
# epilogues as the prefill pipeline uses them: fp32 store (qkv), fp32 residual add (o, down), SwiGLU -> bf16 (gate/up)
for name, N, K, epi in [("qkv", 5120, 3072, 0), ("o", 3072, 3072, 2), ("gate_up", 16384, 3072, 3), ("down", 3072, 8192, 2)]:
    A = rng.integers(0x3C00, 0x3F80, size=(M, K), dtype=np.uint16)
    W = rng.integers(0x3C00, 0x3F80, size=(N, K), dtype=np.uint16)
    _, ms = _capi.op_gemm_bf16(A, W, epilogue=epi, iters=10)
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    print(f"{name:8s} epilogue={epi} M={M} N={N} K={K}: {ms:.3f} ms  {tf:.1f} TFLOP/s  ({tf / peak:.2%} of measured cuBLAS bf16 burst {peak})")

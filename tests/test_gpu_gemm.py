"""tcgen05/TMEM prefill GEMM (gabby_b200/csrc/gemm_tcgen05.cuh) through the C-ABI against a float64
reference on the same bf16 inputs. Products of bf16 values are exact in fp32; accumulation is fp32 on the
tensor cores, so the tolerance is a few fp32 ulps of the row sum (1e-3 relative to the output scale)."""
import numpy as np
import pytest

from gabby_b200 import synth

pytestmark = pytest.mark.gpu


def _ref(A, W):
    return synth.bf16_bits_to_f32(A).astype(np.float64) @ synth.bf16_bits_to_f32(W).astype(np.float64).T


def _rand_bits(rng, shape, scale):
    return synth.f32_to_bf16_bits((rng.standard_normal(shape) * scale).astype(np.float32))


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 512), (200, 384, 2048), (1, 128, 128), (384, 128, 8192)])
def test_gemm_fp32_store_matches_float64(M, N, K):
    from gabby_b200 import _capi
    rng = np.random.default_rng(M + N + K)
    A, W = _rand_bits(rng, (M, K), 1.0), _rand_bits(rng, (N, K), 1.0 / np.sqrt(K))
    C, _ = _capi.op_gemm_bf16(A, W, epilogue=0)
    ref = _ref(A, W)
    assert np.abs(C - ref).max() < 1e-3 * max(1.0, np.abs(ref).max())


def test_gemm_epilogues():
    from gabby_b200 import _capi
    rng = np.random.default_rng(5)
    M, N, K = 160, 256, 256
    A, W = _rand_bits(rng, (M, K), 1.0), _rand_bits(rng, (N, K), 1.0 / np.sqrt(K))
    ref = _ref(A, W)
    Cb, _ = _capi.op_gemm_bf16(A, W, epilogue=1)                       # bf16 store
    assert np.abs(Cb - ref).max() < 2.0 ** -8 * np.abs(ref).max() + 1e-3
    R = rng.standard_normal((M, N)).astype(np.float32)
    Ca, _ = _capi.op_gemm_bf16(A, W, epilogue=2, c_in=R)                # residual add
    assert np.abs(Ca - (R + ref)).max() < 1e-3 * max(1.0, np.abs(ref).max())
    Cs, _ = _capi.op_gemm_bf16(A, W, epilogue=3)                        # SwiGLU over (gate, up) row pairs of W
    g, u = ref[:, 0::2], ref[:, 1::2]
    sw = (g / (1.0 + np.exp(-g))) * u
    assert Cs.shape == (M, N // 2)
    assert np.abs(Cs - sw).max() < 2.0 ** -7 * np.abs(sw).max() + 1e-3


def test_gemm_rejects_bad_shapes():
    from gabby_b200 import _capi
    A, W = np.zeros((128, 96), np.uint16), np.zeros((128, 96), np.uint16)
    with pytest.raises(_capi.B2lError, match="multiple of 64"):
        _capi.op_gemm_bf16(A, W)
    with pytest.raises(_capi.B2lError, match="multiple of 128"):
        _capi.op_gemm_bf16(np.zeros((128, 64), np.uint16), np.zeros((96, 64), np.uint16))

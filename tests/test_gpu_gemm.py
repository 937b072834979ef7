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


# N % 256 == 0 runs the persistent 128 x 256-tile kernel: (2500, 2304, 256) is 180 tiles on 148 CTAs (second tile per CTA,
# both TMEM accumulators, ring wrap across tiles, narrow last column band, ragged last row tile)
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 512), (200, 384, 2048), (1, 128, 128), (384, 128, 8192),
                                   (2500, 2304, 256), (300, 512, 4096), (5000, 2048, 64), (1, 256, 64)])
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


def test_gemm_epilogues_persistent_many_tiles():
    from gabby_b200 import _capi
    rng = np.random.default_rng(11)
    M, N, K = 2000, 2560, 192       # 16 x 10 tiles of 128 x 256 on 148 CTAs
    A, W = _rand_bits(rng, (M, K), 1.0), _rand_bits(rng, (N, K), 1.0 / np.sqrt(K))
    ref = _ref(A, W)
    Cb, _ = _capi.op_gemm_bf16(A, W, epilogue=1)
    assert np.abs(Cb - ref).max() < 2.0 ** -8 * np.abs(ref).max() + 1e-3
    R = rng.standard_normal((M, N)).astype(np.float32)
    Ca, _ = _capi.op_gemm_bf16(A, W, epilogue=2, c_in=R)
    assert np.abs(Ca - (R + ref)).max() < 1e-3 * max(1.0, np.abs(ref).max())
    Cs, _ = _capi.op_gemm_bf16(A, W, epilogue=3)
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


# ---------------------------------------------------------------------------------------------
# the GEMM prefill path inside the engine against the oracle with the same rounding points
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("preset,layers,seed,lens", [
    ("tiny", None, 1234, [70]),
    ("tiny128", None, 77, [33, 90, 5]),          # ragged batch, one sequence shorter than a GEMM tile
    ("1b", 2, 5, [130]),                          # full 1B width
])
def test_gemm_prefill_matches_oracle_with_bf16_activations(preset, layers, seed, lens):
    from oracle import pyoracle as po
    from tests.helpers import synth_tensors, make_engine, contiguous_tables, cosine
    arch, tensors = synth_tensors(preset, layers, seed)
    prompts = [synth.synth_prompt(n, arch.vocab_size, arch.bos_token_id, 300 + i) for i, n in enumerate(lens)]
    eng = make_engine(arch, tensors, max_batch=len(lens), max_positions=256, max_prefill_tokens=256)
    eng.set_prefill_mode(1)
    eng.set_taps(True)
    bt = contiguous_tables(len(lens), eng.max_blocks)
    first = eng.prefill(prompts, [0] * len(lens), bt)
    logits = eng.logits(0, len(lens))
    hidden_last = eng.hidden(arch.num_hidden_layers, 0, sum(lens))
    eng.set_taps(False)
    ids, _ = eng.decode_loop(first, lens, bt, 8)
    om = po.OracleModel(arch, tensors, 256)
    row = 0
    for i, p in enumerate(prompts):
        s = om.seq(po.ORC_KV_BF16 | po.ORC_ACT_BF16 | po.ORC_QP_BF16)   # bf16 linear inputs, bf16 q and softmax probabilities
        ol, oh = s.forward(p, want_hidden=True)
        # bf16 activations on both sides, but the GPU rounds the lm_head input only on the oracle side and sums in a
        # different order: tolerance 6e-2 on logits of scale 2-6 (observed 3e-2), cosine 0.9999
        assert np.abs(logits[i] - ol[0]).max() < 6e-2 and cosine(logits[i], ol[0]) > 0.9999, i
        hd = np.abs(hidden_last[row:row + len(p)] - oh[arch.num_hidden_layers])
        assert hd.max() < 8e-2 and hd.mean() < 1e-2 and cosine(hidden_last[row:row + len(p)], oh[arch.num_hidden_layers]) > 0.9999, (i, float(hd.max()))
        row += len(p)
        s.set_flags(po.ORC_KV_BF16)                       # decode keeps fp32 activations
        tok, want = int(np.argmax(ol[0])), []
        for _ in range(9):
            want.append(tok)
            lg, _ = s.forward([tok])
            tok = int(np.argmax(lg[0]))
        got = [int(first[i])] + ids[:, i].tolist()
        assert got == want, (i, got, want)


def test_gemm_prefill_continues_cached_context_like_the_exact_path():
    from tests.helpers import synth_tensors, make_engine, contiguous_tables
    arch, tensors = synth_tensors("tiny", None, 1234)
    prompt = synth.synth_prompt(150, arch.vocab_size, arch.bos_token_id, 9)
    outs = []
    for mode in (0, 1):
        eng = make_engine(arch, tensors, max_positions=256, max_prefill_tokens=256)
        eng.set_prefill_mode(mode)
        bt = contiguous_tables(1, eng.max_blocks)
        eng.prefill([prompt[:70]], [0], bt)
        nxt = eng.prefill([prompt[70:]], [70], bt)      # second call appends to the cached 70 tokens
        outs.append((int(nxt[0]), eng.logits(0, 1)[0].copy()))
    assert outs[0][0] == outs[1][0]
    assert np.abs(outs[0][1] - outs[1][1]).max() < 5e-2


# ---------------------------------------------------------------------------------------------
# tcgen05 / TMEM flash prefill attention (flash_prefill_tc.cuh)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("preset,layers,lens", [
    ("tiny128", None, [400, 140, 5, 257]),     # head_dim 128: both query tiles of a CTA, 1-4 kv tiles, ragged tails, a 1-row second tile
    ("1b", 2, [300, 129]),                     # head_dim 64, GQA group 4
    ("3b", 2, [260]),                          # head_dim 128, GQA group 3
])
def test_flash_tc_matches_mma_sync_kernel_and_oracle(preset, layers, lens, monkeypatch):
    """Same prompts through the prefill pipeline with the tcgen05 attention kernel and with the mma.sync kernel (B2L_FLASH_TC=0):
    every row of the last layer's residual stream must agree (both round q, K, V and P to bf16 at the same points; the
    running-max schedules differ: 64- vs 128-token tiles), and the tcgen05 run must match the oracle like the old kernel does."""
    from oracle import pyoracle as po
    from tests.helpers import synth_tensors, make_engine, contiguous_tables, cosine
    arch, tensors = synth_tensors(preset, layers, 21)
    prompts = [synth.synth_prompt(n, arch.vocab_size, arch.bos_token_id, 900 + i) for i, n in enumerate(lens)]
    L = arch.num_hidden_layers
    outs = []
    for tc in ("1", "0"):
        monkeypatch.setenv("B2L_FLASH_TC", tc)
        eng = make_engine(arch, tensors, max_batch=len(lens), max_positions=512, max_prefill_tokens=sum(lens) + 8, num_pages=len(lens) * 32)
        eng.set_prefill_mode(1)
        eng.set_taps(True)
        bt = contiguous_tables(len(lens), eng.max_blocks)
        first = eng.prefill(prompts, [0] * len(lens), bt)
        outs.append((first.copy(), eng.logits(0, len(lens)).copy(), eng.hidden(L, 0, sum(lens)).copy()))
        eng.close()
    (f1, lg1, h1), (f0, lg0, h0) = outs
    scale = np.abs(h0).max()
    assert np.abs(h1 - h0).max() < 2e-2 * max(1.0, scale) and cosine(h1, h0) > 0.9999, float(np.abs(h1 - h0).max())
    assert np.abs(lg1 - lg0).max() < 6e-2   # bf16 activations: a value on a rounding boundary may flip (same bound as against the oracle)
    om = po.OracleModel(arch, tensors, 512)
    row = 0
    for i, p in enumerate(prompts):
        s = om.seq(po.ORC_KV_BF16 | po.ORC_ACT_BF16 | po.ORC_QP_BF16)
        ol, oh = s.forward(p, want_hidden=True)
        assert np.abs(lg1[i] - ol[0]).max() < 6e-2 and cosine(lg1[i], ol[0]) > 0.9999, (i, float(np.abs(lg1[i] - ol[0]).max()))
        hd = np.abs(h1[row:row + len(p)] - oh[L])
        assert hd.max() < 8e-2 and hd.mean() < 1e-2, (i, float(hd.max()))
        row += len(p)


def test_flash_tc_continues_cached_context():
    """Second prefill call appends to 70 cached tokens: query positions start mid-page and mid-tile."""
    from tests.helpers import synth_tensors, make_engine, contiguous_tables
    arch, tensors = synth_tensors("tiny128", None, 77)
    prompt = synth.synth_prompt(330, arch.vocab_size, arch.bos_token_id, 9)
    outs = []
    for mode in (0, 1):
        eng = make_engine(arch, tensors, max_positions=512, max_prefill_tokens=512, num_pages=40)
        eng.set_prefill_mode(mode)
        bt = contiguous_tables(1, eng.max_blocks)
        eng.prefill([prompt[:70]], [0], bt)
        nxt = eng.prefill([prompt[70:]], [70], bt)
        outs.append((int(nxt[0]), eng.logits(0, 1)[0].copy()))
        eng.close()
    assert outs[0][0] == outs[1][0]
    assert np.abs(outs[0][1] - outs[1][1]).max() < 5e-2

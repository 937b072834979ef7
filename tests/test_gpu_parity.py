"""GPU parity tests (run on a B200 with -m gpu): the CUDA path through the C-ABI (include/b2l.h)
against the CPU oracle on the same seeded inputs, and against the committed HF golden fixtures.

Tolerances (bf16 weights, bf16 KV cache, fp32 activations and accumulation on both sides; the
only differences are fp32 summation order, exp/rsqrt implementations, and the rare bf16
rounding flip of a cached K/V element):
    logits   max-abs <= LOGIT_ATOL (logit scale ~2-6), cosine >= 0.99999
    hidden   max-abs <= HIDDEN_ATOL
    greedy token ids and argmax indices: bit-identical
"""
import numpy as np
import pytest

from gabby_b200 import synth
from tests.helpers import load_golden, synth_tensors, cosine, make_engine, contiguous_tables

pytestmark = pytest.mark.gpu

LOGIT_ATOL = 4e-3
HIDDEN_ATOL = 4e-3
COS_MIN = 0.99999


def _po():
    from oracle import pyoracle as po
    return po


def _capi():
    from gabby_b200 import _capi
    return _capi


# ---------------------------------------------------------------------------------------------
# single kernels
# ---------------------------------------------------------------------------------------------

def _ref_gemv(Wb, x, norm_b=None, eps=1e-5, mode=0, y_in=None):
    W = synth.bf16_bits_to_f32(Wb).astype(np.float64)
    x = x.astype(np.float64)
    if norm_b is not None:
        inv = 1.0 / np.sqrt((x * x).mean(axis=1, keepdims=True) + eps)
        x = synth.bf16_bits_to_f32(norm_b).astype(np.float64)[None, :] * (x * inv)
    y = x @ W.T
    if mode == 2:
        g, u = y[:, 0::2], y[:, 1::2]
        return (g / (1.0 + np.exp(-g))) * u
    if mode == 1:
        return y_in.astype(np.float64) + y
    return y


@pytest.mark.parametrize("N,K,mode,norm", [
    (3072, 2048, 0, True),     # 1B fused QKV (+ fused input RMSNorm)
    (2048, 2048, 1, False),    # 1B O-proj (+ residual add)
    (4096, 2048, 2, True),     # 1B gate/up slice (+ fused post-attn RMSNorm, SwiGLU)
    (2048, 8192, 1, False),    # 1B down (+ residual add)
    (4096, 2048, 0, True),     # 1B lm_head slice (+ fused final norm)
    (512, 28672, 1, False),    # 70B down: K larger than one shared-memory tile
    (258, 3584, 0, False),     # ragged: N not a multiple of the CTA row block, K = 14 * 256
    (64, 136, 0, False),       # K not a multiple of 256
])
@pytest.mark.parametrize("B", [1, 3, 8])
def test_gemv_against_fp64(N, K, mode, norm, B):
    rng = np.random.default_rng(N * 7 + K + B)
    Wb = synth.f32_to_bf16_bits((rng.standard_normal((N, K)) * (1.0 / np.sqrt(K))).astype(np.float32))
    x = rng.standard_normal((B, K)).astype(np.float32)
    nb = synth.f32_to_bf16_bits((1.0 + 0.1 * rng.standard_normal(K)).astype(np.float32)) if norm else None
    y0 = rng.standard_normal((B, N)).astype(np.float32) if mode == 1 else None
    y, _ = _capi().op_gemv(Wb, x, y_in=y0, norm_w_bits=nb, mode=mode)
    ref = _ref_gemv(Wb, x, nb, 1e-5, mode, y0)
    assert np.abs(y - ref).max() < 2e-4 * max(1.0, np.abs(ref).max())


def test_argmax_first_max_and_edge_values():
    x = np.full((4, 128256), -1.0, np.float32)
    x[0, 77] = 3.0; x[0, 90000] = 3.0            # tie -> lowest index
    x[1, 128255] = 0.5                           # last element
    x[2, :] = -np.inf; x[2, 5] = -1e30           # -inf everywhere else
    x[3, 0] = -1.0                               # all equal -> index 0
    out = _capi().op_argmax(x)
    assert out.tolist() == [77, 128255, 5, 0]
    rng = np.random.default_rng(0)
    y = rng.standard_normal((8, 50001)).astype(np.float32)
    assert np.array_equal(_capi().op_argmax(y), y.argmax(axis=1).astype(np.int32))
    # fewer columns than the cluster has CTAs (empty slices), and a row length that is not a multiple of anything
    for n in (1, 5, 1031):
        z = rng.standard_normal((3, n)).astype(np.float32)
        assert np.array_equal(_capi().op_argmax(z), z.argmax(axis=1).astype(np.int32)), n


# ---------------------------------------------------------------------------------------------
# whole forward, tiny shapes, oracle + HF golden
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["tiny_s1234", "tiny128_s77", "w1b_l2_s5"])
def test_prefill_and_greedy_match_oracle_and_hf_golden(name):
    po = _po()
    g, arch, seed = load_golden(name)
    layers = int(g["layers"])
    _, tensors = synth_tensors(str(g["preset"]), None if layers < 0 else layers, seed)
    prompt = g["prompt"]
    n_new = len(g["greedy_ids"])
    eng = make_engine(arch, tensors, max_positions=128, max_prefill_tokens=64)
    eng.set_taps(True)
    bt = contiguous_tables(1, eng.max_blocks)
    first = eng.prefill([prompt], [0], bt)
    om = po.OracleModel(arch, tensors, 128)
    os_ = om.seq(po.ORC_KV_BF16)
    ologits, ohidden = os_.forward(prompt, logits_all=True, want_hidden=True)
    # last-token logits vs oracle and vs HF
    lg = eng.logits(0, 1)[0]
    assert np.abs(lg - ologits[-1]).max() < LOGIT_ATOL
    assert cosine(lg, ologits[-1]) > COS_MIN
    assert np.abs(lg[g["logit_cols"]] - g["logits"][-1]).max() < 3e-2      # HF keeps fp32 KV: looser
    # residual stream after every layer, every prompt position
    L = arch.num_hidden_layers
    for slab in range(L + 2):
        hid = eng.hidden(slab, 0, len(prompt))
        assert np.abs(hid - ohidden[slab]).max() < HIDDEN_ATOL, slab
    assert first[0] == g["greedy_ids"][0] == po.lib().orc_argmax(ologits[-1].ctypes.data, arch.vocab_size)
    # greedy continuation: step API, bit-identical ids vs HF golden and oracle
    eng.set_taps(False)
    ids = [int(first[0])]
    pos = len(prompt)
    for i in range(n_new - 1):
        nxt = eng.decode([ids[-1]], [pos], bt)
        step_lg = eng.logits(0, 1)[0]
        ol, _ = os_.forward([ids[-1]])
        assert np.abs(step_lg - ol[0]).max() < LOGIT_ATOL, i
        ids.append(int(nxt[0]))
        pos += 1
    assert ids == g["greedy_ids"].tolist()


def test_decode_loop_equals_stepwise_and_oracle():
    po = _po()
    arch, tensors = synth_tensors("tiny", None, 1234)
    prompt = synth.synth_prompt(21, arch.vocab_size, arch.bos_token_id, 11)
    eng = make_engine(arch, tensors, max_positions=128)
    bt = contiguous_tables(1, eng.max_blocks)
    first = eng.prefill([prompt], [0], bt)
    loop_ids, ms = eng.decode_loop(first, [len(prompt)], bt, 40)
    assert ms > 0
    eng2 = make_engine(arch, tensors, max_positions=128)
    first2 = eng2.prefill([prompt], [0], bt)
    step_ids, tok, pos = [], first2, len(prompt)
    for _ in range(40):
        tok = eng2.decode(tok, [pos], bt)
        step_ids.append(int(tok[0]))
        pos += 1
    assert loop_ids[:, 0].tolist() == step_ids
    oids, margins = po.OracleModel(arch, tensors, 128).seq(po.ORC_KV_BF16).greedy(prompt, 41)
    assert int(first[0]) == int(oids[0])
    assert loop_ids[:, 0].tolist() == oids[1:].tolist(), f"min oracle margin {margins.min()}"


def test_paged_kv_is_permutation_invariant_and_pages_hold_bf16_kv():
    arch, tensors = synth_tensors("tiny", None, 1234)
    prompt = synth.synth_prompt(37, arch.vocab_size, arch.bos_token_id, 5)
    outs = []
    for perm_seed in (None, 3):
        eng = make_engine(arch, tensors, max_positions=128, num_pages=40)
        bt = np.arange(eng.max_blocks, dtype=np.int32)[None, :]
        if perm_seed is not None:
            bt = np.random.default_rng(perm_seed).permutation(40)[: eng.max_blocks].astype(np.int32)[None, :]
        first = eng.prefill([prompt], [0], bt)
        ids, _ = eng.decode_loop(first, [len(prompt)], bt, 12)
        outs.append((int(first[0]), ids[:, 0].tolist(), eng.logits(0, 1).copy(), eng.kv_page(1, int(bt[0, 1]), 0).copy(),
                     eng.kv_page(1, int(bt[0, 1]), 1).copy()))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    assert np.array_equal(outs[0][2], outs[1][2])          # bit-identical logits
    assert np.array_equal(outs[0][3], outs[1][3]) and np.array_equal(outs[0][4], outs[1][4])
    assert outs[0][3].any() and outs[0][4].any()


def test_batched_decode_ragged_lengths_matches_per_sequence_oracle():
    po = _po()
    arch, tensors = synth_tensors("tiny128", None, 77)
    lens = [5, 33, 17, 1, 48]
    prompts = [synth.synth_prompt(n, arch.vocab_size, arch.bos_token_id, 100 + i) for i, n in enumerate(lens)]
    eng = make_engine(arch, tensors, max_batch=5, max_positions=96, max_prefill_tokens=128)
    bt = contiguous_tables(5, eng.max_blocks)
    first = eng.prefill(prompts, [0] * 5, bt)
    ids, _ = eng.decode_loop(first, lens, bt, 10)
    om = po.OracleModel(arch, tensors, 96)
    for i, p in enumerate(prompts):
        oids, margins = om.seq(po.ORC_KV_BF16).greedy(p, 11)
        assert int(first[i]) == int(oids[0]), i
        assert ids[:, i].tolist() == oids[1:].tolist(), (i, float(margins.min()))


@pytest.mark.parametrize("n_seq", [5, 12, 20, 40])
def test_batched_decode_on_tensor_cores_matches_per_sequence_oracle(n_seq):
    """2..64 rows per step: projections run as tcgen05 skinny GEMMs (hi/lo bf16 split of the fp32 activations); 40 rows =
    two row groups (32 + 8): the fused QKV-reduce + RoPE + KV-append epilogue with a row offset."""
    po = _po()
    arch, tensors = synth_tensors("1b", 2, 5)
    lens = [(7 * i + 3) % 40 + 1 for i in range(n_seq)]
    prompts = [synth.synth_prompt(n, arch.vocab_size, arch.bos_token_id, 300 + i) for i, n in enumerate(lens)]
    eng = make_engine(arch, tensors, max_batch=n_seq, max_positions=64, max_prefill_tokens=64 * n_seq)
    assert eng.info().batched_tensor_core == 1
    eng.set_prefill_mode(0)
    bt = contiguous_tables(n_seq, eng.max_blocks)
    first = eng.prefill(prompts, [0] * n_seq, bt)
    ids, _ = eng.decode_loop(first, lens, bt, 6)
    last_logits = eng.logits(0, n_seq).copy()
    om = po.OracleModel(arch, tensors, 64)
    for i, p in enumerate(prompts):
        seq = om.seq(po.ORC_KV_BF16)
        oids, margins = seq.greedy(p, 7)
        assert int(first[i]) == int(oids[0]), i
        assert ids[:, i].tolist() == oids[1:].tolist(), (i, float(margins.min()))
        # logits of the last decode step against the oracle's (fp32 activations on both sides)
        seq.reset()
        ol, _ = seq.forward(np.concatenate([p, oids[:6]]).astype(np.int32))
        assert np.abs(last_logits[i] - ol[0]).max() < 5e-3, (i, float(np.abs(last_logits[i] - ol[0]).max()))
    eng.close()


def test_chunked_prefill_continues_a_cached_sequence():
    arch, tensors = synth_tensors("tiny", None, 1234)
    prompt = synth.synth_prompt(30, arch.vocab_size, arch.bos_token_id, 8)
    a = make_engine(arch, tensors, max_positions=64)
    bt = contiguous_tables(1, a.max_blocks)
    one = a.prefill([prompt], [0], bt)
    la = a.logits(0, 1).copy()
    b = make_engine(arch, tensors, max_positions=64)
    b.prefill([prompt[:13]], [0], bt)
    two = b.prefill([prompt[13:]], [13], bt)
    assert one[0] == two[0]
    assert np.abs(la - b.logits(0, 1)).max() < 1e-4


def test_on_device_synthetic_weights_are_bit_identical_to_uploaded_ones():
    arch, tensors = synth_tensors("tiny128", None, 77)
    prompt = synth.synth_prompt(9, arch.vocab_size, arch.bos_token_id, 2)
    a = make_engine(arch, tensors, max_positions=64)
    b = make_engine(arch, None, max_positions=64, synth_seed=77)
    bt = contiguous_tables(1, a.max_blocks)
    assert a.prefill([prompt], [0], bt)[0] == b.prefill([prompt], [0], bt)[0]
    assert np.array_equal(a.logits(0, 1), b.logits(0, 1))


def test_error_behaviour_mirrors_reference_conventions():
    """Errors come back as non-zero + message (the C++ adapter rethrows std::runtime_error, which
    gabby's server maps to HTTP 500: /root/reference/src/http/server.cc:371-378)."""
    capi = _capi()
    arch, tensors = synth_tensors("tiny", None, 1234)
    eng = make_engine(arch, tensors, max_positions=64, num_pages=4)
    bt = contiguous_tables(1, eng.max_blocks)
    with pytest.raises(capi.B2lError, match="token id out of range"):
        eng.prefill([[arch.vocab_size]], [0], bt)
    with pytest.raises(capi.B2lError, match="max_positions"):
        eng.prefill([np.zeros(65, np.int32)], [0], bt)
    with pytest.raises(capi.B2lError, match="page id outside the pool"):
        eng.prefill([np.zeros(40, np.int32)], [0], np.full((1, eng.max_blocks), 9, np.int32))
    with pytest.raises(capi.B2lError, match="n_seq out of range"):
        eng.decode([1, 2], [0, 0], contiguous_tables(2, eng.max_blocks))
    e2 = capi.Engine(arch, _po().rope_table(arch, 64), max_positions=64)
    with pytest.raises(capi.B2lError, match="unknown tensor name"):
        e2.upload("model.layers.0.bogus", np.zeros(4, np.uint16), (4,))
    with pytest.raises(capi.B2lError, match="shape mismatch"):
        e2.upload("model.norm.weight", np.zeros(4, np.uint16), (4,))
    with pytest.raises(capi.B2lError, match="missing"):
        e2.finalize()
    with pytest.raises(capi.B2lError, match="finalize has not been called"):
        e2.decode([1], [0], bt)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[0]: full Llama-3.2-1B architecture, 32-token prompt + 32 greedy tokens
# ---------------------------------------------------------------------------------------------

def test_config1_llama32_1b_32_prompt_32_greedy_matches_cpu_oracle():
    po = _po()
    arch = synth.preset("1b")
    seed = 20261018
    specs = synth.tensor_specs(arch)
    tensors = {n: synth.gen_tensor_bits(n, int(np.prod(s)), sc, off, seed) for n, s, sc, off in specs}
    prompt = synth.synth_prompt(32, arch.vocab_size, arch.bos_token_id, seed + 1)
    eng = make_engine(arch, tensors, max_positions=128, max_prefill_tokens=64)
    bt = contiguous_tables(1, eng.max_blocks)
    first = eng.prefill([prompt], [0], bt)
    lg0 = eng.logits(0, 1)[0].copy()
    ids, ms = eng.decode_loop(first, [32], bt, 31)
    gpu_ids = [int(first[0])] + ids[:, 0].tolist()
    om = po.OracleModel(arch, tensors, 128)
    s = om.seq(po.ORC_KV_BF16)
    ol, _ = s.forward(prompt)
    assert np.abs(lg0 - ol[0]).max() < LOGIT_ATOL and cosine(lg0, ol[0]) > COS_MIN
    oids, margins = om.seq(po.ORC_KV_BF16).greedy(prompt, 32)
    assert gpu_ids == oids.tolist(), f"min oracle top-1 margin {float(margins.min()):.4g}"
    assert len(set(gpu_ids)) > 4          # non-degenerate continuation


# ---------------------------------------------------------------------------------------------
# persistent megakernel (decode mode 1) against the multi-kernel path (mode 0) and the oracle
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("preset,layers,seed,n_prompt,n_new", [
    ("tiny", None, 1234, 21, 24),        # hd 32, group 4, H 256 (m=1), I 512 (m=2)
    ("tiny128", None, 77, 40, 100),      # hd 128, group 2, untied head; context crosses 64 and 128 (1 -> 2 -> 3 splits)
    ("1b", 2, 5, 12, 8),                 # full 1B width: K 2048 (ks 1) and K 8192 (ks 4), V 128256
])
def test_megakernel_matches_multikernel_path_and_oracle(preset, layers, seed, n_prompt, n_new):
    po = _po()
    arch, tensors = synth_tensors(preset, layers, seed)
    prompt = synth.synth_prompt(n_prompt, arch.vocab_size, arch.bos_token_id, seed + 3)
    res = {}
    for mode in (0, 1):
        eng = make_engine(arch, tensors, max_positions=256)
        assert eng.info().decode_mode == 1, "megakernel should be available for this shape"
        eng.set_decode_mode(mode)
        bt = contiguous_tables(1, eng.max_blocks)
        first = eng.prefill([prompt], [0], bt)
        ids, ms = eng.decode_loop(first, [n_prompt], bt, n_new)
        # one more token through the per-step call, continuing from the loop's state
        nxt = eng.decode([ids[-1, 0]], [n_prompt + n_new], bt)
        res[mode] = (int(first[0]), ids[:, 0].tolist(), int(nxt[0]), eng.logits(0, 1)[0].copy(), eng.info().kernels_launched)
        eng.close()
    assert res[0][0] == res[1][0]
    assert res[0][1] == res[1][1], "greedy ids differ between megakernel and multi-kernel path"
    assert res[0][2] == res[1][2]
    assert np.abs(res[0][3] - res[1][3]).max() < 1e-3      # fp32 summation order + rare bf16 KV rounding flips
    assert res[1][4] < res[0][4]          # far fewer launches
    oids, margins = po.OracleModel(arch, tensors, 256).seq(po.ORC_KV_BF16).greedy(prompt, n_new + 2)
    assert [res[1][0]] + res[1][1] + [res[1][2]] == oids.tolist(), f"min oracle margin {float(margins.min()):.3g}"


def test_megakernel_relaunch_state_is_clean():
    """Barrier epoch, argmax keys and split counters must carry over correctly between launches."""
    arch, tensors = synth_tensors("tiny", None, 1234)
    prompt = synth.synth_prompt(70, arch.vocab_size, arch.bos_token_id, 4)
    eng = make_engine(arch, tensors, max_positions=256)
    bt = contiguous_tables(1, eng.max_blocks)
    first = eng.prefill([prompt], [0], bt)
    a, _ = eng.decode_loop(first, [70], bt, 50)
    b, _ = eng.decode_loop(first, [70], bt, 50)          # same start: identical continuation
    assert a[:, 0].tolist() == b[:, 0].tolist()
    ids, tok, pos = [], first, 70
    for _ in range(50):                                  # 50 single-step launches
        tok = eng.decode(tok, [pos], bt); pos += 1
        ids.append(int(tok[0]))
    assert ids == a[:, 0].tolist()

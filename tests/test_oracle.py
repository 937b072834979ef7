"""CPU gate: the oracle (oracle/llama_oracle.cc) against the committed HF golden fixtures.

The reference holds no golden vectors for the forward pass, so these
fixtures (tests/golden/make_golden.py, HF transformers 5.5.0 fp32 CPU) are what pins it.
"""
import os
import tempfile

import numpy as np
import pytest

from gabby_b200 import synth
from oracle import pyoracle as po
from tests.helpers import load_golden, synth_tensors, cosine

# w3b / w8b / w70b: two-layer variants at the Llama-3.2-3B (GQA group 3), Llama-3.1-8B (rope factor 8, untied head) and
# Llama-3.1-70B (GQA group 8) widths -- the shapes of BASELINE configs[2..4]
CASES = ["tiny_s1234", "tiny128_s77", "w1b_l2_s5", "w3b_l2_s34", "w8b_l2_s9", "w70b_l2_s71"]   
LOGIT_ATOL = 5e-5   # fp32 vs fp32, different summation order
HIDDEN_ATOL = 5e-5


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_hf_golden(name):
    g, arch, seed = load_golden(name)
    layers = int(g["layers"])
    _, tensors = synth_tensors(str(g["preset"]), None if layers < 0 else layers, seed)
    m = po.OracleModel(arch, tensors, 128)
    s = m.seq(0)
    logits, hidden = s.forward(g["prompt"], logits_all=True, want_hidden=True)
    assert np.abs(logits[:, g["logit_cols"]] - g["logits"]).max() < LOGIT_ATOL
    assert np.array_equal(np.argsort(-logits, axis=1)[:, :1], g["top8"][:, :1])
    hf = g["hidden"]  # [L+1, n, cols]: embeddings, layer outputs 1..L-1, then final norm
    L = arch.num_hidden_layers
    for i in range(L):
        assert np.abs(hidden[i][:, g["hidden_cols"]] - hf[i]).max() < HIDDEN_ATOL, i
    assert np.abs(hidden[L + 1][:, g["hidden_cols"]] - hf[L]).max() < HIDDEN_ATOL
    # greedy decode through the KV cache: bit-identical ids, per-step logits within tolerance
    s2 = m.seq(0)
    ids, margins = s2.greedy(g["prompt"], len(g["greedy_ids"]))
    assert np.array_equal(ids, g["greedy_ids"])
    assert margins.min() > 0
    s3 = m.seq(0)
    s3.forward(g["prompt"])
    for i, tok in enumerate(g["greedy_ids"][:-1]):
        lg, _ = s3.forward([tok])
        assert np.abs(lg[0, g["logit_cols"]] - g["step_logits"][i]).max() < LOGIT_ATOL


def test_rope_inv_freq_matches_hf():
    for name in CASES:
        g, arch, _ = load_golden(name)
        assert np.abs(po.rope_inv_freq(arch) - g["inv_freq"]).max() < 1e-7
    # 1B/8B published scaling parameters: low band divided by factor, high band untouched
    a = synth.preset("1b")
    f = po.rope_inv_freq(a)
    assert f[0] == 1.0 and abs(f[-1] * 32.0 - 1.0 / 500000.0 ** (62 / 64)) < 1e-12


def test_rope_table_definition():
    a = synth.preset("tiny")
    t = po.rope_table(a, 50)
    inv = po.rope_inv_freq(a)
    ang = (np.arange(50, dtype=np.float32)[:, None] * inv[None, :]).astype(np.float32)
    assert np.array_equal(t[..., 0], np.cos(ang.astype(np.float64)).astype(np.float32))
    assert np.array_equal(t[..., 1], np.sin(ang.astype(np.float64)).astype(np.float32))


def test_synth_generator_restatement_bit_exact():
    for name, n, scale, off in [("model.embed_tokens.weight", 100003, 0.05, 0.0),
                                ("model.layers.3.input_layernorm.weight", 2048, 0.1, 1.0),
                                ("x", 1 << 16, float(np.sqrt(3.0 / 8192)), 0.0)]:
        ts = synth.tensor_seed(name, 99)
        assert np.array_equal(synth.gen_tensor_bits(name, n, scale, off, 99), po.synth_tensor(ts, n, scale, off))


def test_bf16_round_trip_and_rne():
    x = np.array([1.0, 1.00390625, 1.005859375, -3.14159, 65504.0, 1e-30], dtype=np.float32)
    b = synth.f32_to_bf16_bits(x)
    assert b[0] == 0x3F80
    assert b[1] == 0x3F80          # tie -> even
    assert b[2] == 0x3F81
    y = synth.bf16_bits_to_f32(b)
    assert np.all(np.abs(y - x) <= np.abs(x) * 2.0 ** -8)


def test_chunked_prefill_equals_one_shot_and_flags():
    arch, tensors = synth_tensors("tiny", None, 1234)
    m = po.OracleModel(arch, tensors, 64)
    prompt = synth.synth_prompt(20, arch.vocab_size, arch.bos_token_id, 3)
    a = m.seq(0)
    la, _ = a.forward(prompt)
    b = m.seq(0)
    b.forward(prompt[:7])
    b.forward(prompt[7:8])
    lb, _ = b.forward(prompt[8:])
    assert np.abs(la - lb).max() < 2e-5
    # bf16 rounding points move logits a little, not a lot
    c = m.seq(po.ORC_KV_BF16 | po.ORC_ACT_BF16)
    lc, _ = c.forward(prompt)
    assert 1e-6 < np.abs(la - lc).max() < 0.1
    assert cosine(la, lc) > 0.999


def test_capacity_and_bad_token_errors():
    arch, tensors = synth_tensors("tiny", None, 1234)
    m = po.OracleModel(arch, tensors, 8)
    s = m.seq(0)
    with pytest.raises(RuntimeError):
        s.forward(np.zeros(9, np.int32))
    with pytest.raises(RuntimeError):
        s.forward([arch.vocab_size])
    with pytest.raises(ValueError):
        po.OracleModel(arch, {k: v for k, v in list(tensors.items())[:-1]}, 8)


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(po.__file__), "_ref", "liboracle_ref.so"))
                    and not os.path.isdir("/root/reference/src"), reason="oracle/_ref not built and no reference here")
def test_loading_through_reference_parsers_gives_same_logits():
    """oracle/_ref: gabby's LoadConfig + Safetensors + json parser feed the same forward."""
    arch, tensors = synth_tensors("tiny", None, 1234)
    with tempfile.TemporaryDirectory() as d:
        synth.write_model_dir(d, arch, 1234)
        mr = po.OracleModel.from_dir_via_reference(d, arch, 64)
        md = po.OracleModel.from_dir(d, arch, 64)
        mm = po.OracleModel(arch, tensors, 64)
        prompt = synth.synth_prompt(9, arch.vocab_size, arch.bos_token_id, 3)
        lr, _ = mr.seq(0).forward(prompt)
        ld, _ = md.seq(0).forward(prompt)
        lm, _ = mm.seq(0).forward(prompt)
        assert np.array_equal(lr, lm) and np.array_equal(ld, lm)
        # the reference's own Generate is a constant-string stub (generator.cc:33-38)
        import ctypes as C
        buf = C.create_string_buffer(128)
        r = po.ref_lib()
        h = r.orc_ref_load(d.encode(), 16)
        r.orc_ref_stub_generate(h, buf, 128)
        assert buf.value == b"hey this is gabby, how are u"
        r.orc_ref_free(h)
        mr.close()


def test_bench_parity_golden_ids_are_what_the_oracle_produces():
    """tests/golden/bench_parity.json (read by bench.py's parity_check legs) is regenerated here for its shortest leg and
    compared, so that the committed ids cannot drift from the oracle (generator: tests/golden/make_bench_parity.py)."""
    import importlib.util
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_bench_parity", os.path.join(path, "make_bench_parity.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    with open(os.path.join(path, "bench_parity.json")) as f:
        gold = json.load(f)
    assert gold["seed"] == mk.SEED
    leg = gold["legs"]["8b_l2"]
    arch, tensors = mk.model(leg["preset"])
    om = po.OracleModel(arch, tensors, 100)
    for i, (n, sd) in enumerate(zip(leg["prompt_lens"], leg["prompt_seeds"])):
        prompt = synth.synth_prompt(n, arch.vocab_size, arch.bos_token_id, sd)
        ids, margins = mk.greedy(om, prompt, len(leg["ids"][i]), po.ORC_KV_BF16)
        assert ids == leg["ids"][i]
        assert np.allclose(margins, leg["margins"][i], atol=1e-5)

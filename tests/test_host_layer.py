"""CPU tests of the host C++ layer (gabby_b200/host/) through its C view: the mirror of gabby's
inference:: interfaces. Edge cases follow what the reference's own tests and SURVEY.md 2.2 name."""
import ctypes as C
import json
import os
import re
import struct
import tempfile

import numpy as np
import pytest

from gabby_b200 import _host, build, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    build.build_all()


def test_host_header_symbols_all_exported():
    src = open(os.path.join(ROOT, "include", "gabby_b200_host.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    syms = sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", src)))
    assert sorted(_host.SYMBOLS) == syms
    L = _host.lib()
    for s in syms:
        assert hasattr(L, s), s


def test_params_from_config_json_llama32_1b_values():
    cfg = synth.hf_config(synth.preset("1b"))
    p = _host.params_from_json(json.dumps(cfg), json.dumps({"eos_token_id": [128001, 128008, 128009]}))
    assert (p.hidden_size, p.intermediate_size, p.num_hidden_layers) == (2048, 8192, 16)
    assert (p.num_attention_heads, p.num_key_value_heads, p.head_dim, p.vocab_size) == (32, 8, 64, 128256)
    assert p.tie_word_embeddings == 1 and p.rope_llama3 == 1 and p.rope_factor == 32.0
    assert abs(p.rms_norm_eps - 1e-5) < 1e-12 and p.rope_theta == 500000.0      # "1e-05" parses (reference parser quirk)
    assert list(p.eos_token_ids[: p.n_eos]) == [128001, 128008, 128009]


def test_params_errors_name_the_key():
    cfg = synth.hf_config(synth.preset("tiny"))
    del cfg["hidden_size"]
    with pytest.raises(_host.HostError, match="hidden_size"):
        _host.params_from_json(json.dumps(cfg))
    cfg = synth.hf_config(synth.preset("tiny"))
    cfg["head_dim"] = 48
    with pytest.raises(_host.HostError, match="head_dim"):
        _host.params_from_json(json.dumps(cfg))
    cfg = synth.hf_config(synth.preset("tiny"))
    cfg["rope_scaling"]["rope_type"] = "yarn"
    with pytest.raises(_host.HostError, match="rope_scaling"):
        _host.params_from_json(json.dumps(cfg))
    with pytest.raises(_host.HostError, match="json"):
        _host.params_from_json("{ not json")


def test_json_string_escapes_are_real():
    """The reference drops the backslash (parser.cc:112-121); tokenizer vocab needs real escapes."""
    tok = {"model": {"type": "BPE", "vocab": {"a": 0, "Ġ": 1, "\n": 2, "\\": 3, "\"": 4, "é": 5}, "merges": []}, "added_tokens": []}
    t = _host.Tokenizer(json.dumps(tok))            # json.dumps writes Ġ, \n, \\, \" escapes
    assert t.detokenize([0]) == "a"
    # U+0120 is the byte-level image of the space byte
    assert t.detokenize([1]) == " "


def test_safetensors_accessor_single_and_sharded_and_errors():
    arch = synth.preset("tiny")
    with tempfile.TemporaryDirectory() as d1, tempfile.TemporaryDirectory() as d3:
        synth.write_model_dir(d1, arch, 9)
        synth.write_model_dir(d3, arch, 9, shards=3)
        n1, f1 = _host.checkpoint_info(d1)
        n3, f3 = _host.checkpoint_info(d3)
        assert (n1, f1) == (len(synth.tensor_specs(arch)), 1) and (n3, f3) == (n1, 3)
        for name, shape, scale, off in synth.tensor_specs(arch)[:5] + synth.tensor_specs(arch)[-2:]:
            bits = synth.gen_tensor_bits(name, int(np.prod(shape)), scale, off, 9)
            h = 0xcbf29ce484222325
            for b in bits.tobytes():
                h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
            for d in (d1, d3):
                s, dt, nb, fnv = _host.checkpoint_tensor(d, name)
                assert s == tuple(shape) and dt == "BF16" and nb == bits.nbytes and fnv == h, (d, name)
        with pytest.raises(_host.HostError, match="no tensor named"):
            _host.checkpoint_tensor(d1, "model.layers.99.bogus")
        # header length beyond the file (the reference's int-shift header read would not notice: safetensors.cc:25-27)
        bad = os.path.join(d1, "model.safetensors")
        with open(bad, "r+b") as f:
            f.write(struct.pack("<Q", 1 << 40))
        with pytest.raises(_host.HostError, match="header length"):
            _host.checkpoint_info(d1)
    with tempfile.TemporaryDirectory() as empty:
        with pytest.raises(_host.HostError, match="model.safetensors"):
            _host.checkpoint_info(empty)


def test_kv_page_allocator_grow_free_exhaust():
    kv = _host.KvAllocator(num_pages=10, page_size=16, max_blocks=6)
    a, b = kv.new_sequence(), kv.new_sequence()
    kv.reserve(a, 1)
    assert kv.table(a).tolist() == [0] and kv.free_pages() == 9
    kv.reserve(a, 16)                                   # still one page
    assert kv.table(a).tolist() == [0]
    kv.reserve(a, 17)
    kv.reserve(b, 40)
    assert kv.table(a).tolist() == [0, 1] and kv.table(b).tolist() == [2, 3, 4] and kv.free_pages() == 5
    with pytest.raises(_host.HostError, match="per-sequence limit"):
        kv.reserve(a, 16 * 6 + 1)
    c = kv.new_sequence()
    with pytest.raises(_host.HostError, match="KV pool exhausted"):
        kv.reserve(c, 16 * 6)                           # needs 6, only 5 free
    assert kv.free_pages() == 5 and kv.table(c).size == 0  # failed reserve changes nothing
    kv.release(a)
    assert kv.free_pages() == 7
    kv.reserve(c, 16 * 6)
    assert sorted(kv.table(c).tolist()) == [0, 1, 5, 6, 7, 8] and kv.free_pages() == 1
    with pytest.raises(_host.HostError, match="unknown sequence"):
        kv.reserve(12345, 1)


def test_tokenizer_reference_contract_and_bpe_against_hf_tokenizers():
    # the reference's three tokenizer tests pin exactly this (tokenizer_test.cc:9-25)
    assert _host.Tokenizer("").tokenize("") == []
    # byte fallback when tokenizer.json has no vocabulary (synthetic dirs)
    t = _host.Tokenizer(json.dumps({"model": {"type": "BPE", "vocab": {}, "merges": []}, "added_tokens": []}))
    assert t.tokenize("hi!") == [104, 105, 33] and t.detokenize([104, 105, 33]) == "hi!"
    # a small byte-level BPE trained with the installed `tokenizers`, compared token by token
    tokenizers = pytest.importorskip("tokenizers")
    from tokenizers import Tokenizer as HfTok, models, pre_tokenizers, decoders, trainers
    hf = HfTok(models.BPE())
    hf.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=True)
    hf.decoder = decoders.ByteLevel()
    corpus = ["the quick brown fox jumps over the lazy dog", "hello world hello there", "attention is all you need",
              "paged kv cache and greedy sampling", "the cat sat on the mat with the hat"] * 20
    hf.train_from_iterator(corpus, trainers.BpeTrainer(vocab_size=400, special_tokens=["<|begin_of_text|>", "<|eot_id|>"],
                                                       initial_alphabet=pre_tokenizers.ByteLevel.alphabet()))
    ours = _host.Tokenizer(hf.to_str())
    for text in ["hello world", "the quick brown fox", " the lazy dog sat", "attention is all", "greedy cache hello"]:
        assert ours.tokenize(text) == hf.encode(text, add_special_tokens=False).ids, text
        assert ours.detokenize(ours.tokenize(text)) == text
    ids = ours.chat_prompt("be brief", "hello")
    assert ids[0] == hf.token_to_id("<|begin_of_text|>") and ids.count(hf.token_to_id("<|eot_id|>")) == 2


LLAMA3_SPLIT = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+")


def _llama3_style_tokenizer():
    """A byte-level BPE trained here, configured exactly like Llama-3's tokenizer.json: Split(regex, isolated) +
    ByteLevel(use_regex=False), ignore_merges, special tokens as added tokens."""
    from tokenizers import Regex, Tokenizer as HfTok, decoders, models, pre_tokenizers, trainers
    hf = HfTok(models.BPE(ignore_merges=True))
    hf.pre_tokenizer = pre_tokenizers.Sequence([
        pre_tokenizers.Split(Regex(LLAMA3_SPLIT), behavior="isolated", invert=False),
        pre_tokenizers.ByteLevel(add_prefix_space=False, trim_offsets=True, use_regex=False)])
    hf.decoder = decoders.ByteLevel()
    corpus = ["The quick brown fox doesn't jump over 12345 lazy dogs!", "I'll say: we've been here, haven't we?  Yes.",
              "naïve café über straße", "日本語のテキストと中文文本", "prices: $1,234.56 or 99% off...", "tabs\tand\nnewlines\r\n  spaces   ",
              "emoji 🙂🙂 and ünïcödé", "HE'S SHOUTING AND SHE'LL HEAR", "x = y**2 + 3*z; // comment"] * 30
    hf.train_from_iterator(corpus, trainers.BpeTrainer(
        vocab_size=600, special_tokens=["<|begin_of_text|>", "<|start_header_id|>", "<|end_header_id|>", "<|eot_id|>"],
        initial_alphabet=pre_tokenizers.ByteLevel.alphabet()))
    return hf


def test_tokenizer_llama3_pretokenizer_matches_hf_tokenizers_on_unicode_digits_contractions_whitespace():
    """SURVEY.md section 8(f) row 1 / the reference's next TODO (tokenizer.cc:6-11): ids must equal what HF `tokenizers` produces
    with Llama-3's own pre-tokenizer configuration."""
    pytest.importorskip("tokenizers")
    hf = _llama3_style_tokenizer()
    ours = _host.Tokenizer(hf.to_str())
    texts = [
        "Hello world", " 123", "12345678", "a 1234 b", "I'll we've doesn't he'd I'm you're IT'S O'Reilly",
        "don't  stop", "x  y   z", "trailing   ", "  leading", "line1\nline2", "a\n\n\nb", "a \n b", "tab\there", "crlf\r\nnext",
        "   \n   x", "wow!!!\n\nnext", "a.b,c;d", " ...!?", "$1,234.56", "99% off", "naïve café", "über straße", "Ünïcödé",
        "日本語のテキスト", "中文 文本 123", "emoji 🙂🙂!", "mixed日本123abc", "ǅ Ⅻ ² ½", "\u00a0nbsp\u2003emspace", "don't<|eot_id|>stop",
        "<|begin_of_text|>hi<|eot_id|>", "a<|start_header_id|>user<|end_header_id|>\n\nhey", "'", "''s", "'S", " 's", "1 2 3",
        "x\u2028y", "snake_case_name __init__", "C++ && C#", "http://example.com/a?b=c&d=1", "",
    ]
    for text in texts:
        want = hf.encode(text, add_special_tokens=False).ids
        got = ours.tokenize(text)
        assert got == want, (text, got, want, [hf.id_to_token(i) for i in want])
    rng = np.random.default_rng(7)
    alphabet = list("abcXYZ 019'\n\t.,!-_é日本🙂 \r") + ["'s", "'ll", "  ", "<|eot_id|>"]
    for _ in range(300):
        text = "".join(rng.choice(alphabet, size=int(rng.integers(1, 24))))
        assert ours.tokenize(text) == hf.encode(text, add_special_tokens=False).ids, repr(text)
    for text in ["I'll say: we've been here", "日本語 and café 🙂", "  spaces\n\nand\ttabs  "]:
        assert ours.detokenize(ours.tokenize(text)) == text


def test_host_argmax_first_max():
    assert _host.argmax([1.0, 3.0, 3.0, -1.0]) == 1
    assert _host.argmax([-np.inf, -np.inf]) == 0


def test_host_rope_table_equals_oracle_table():
    from oracle import pyoracle as po
    for name in ("tiny", "1b", "8b"):
        a = synth.preset(name)
        assert np.array_equal(_host.rope_table(a, 200), po.rope_table(a, 200))


def test_generator_load_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with tempfile.TemporaryDirectory() as d:
        synth.write_model_dir(d, synth.preset("tiny"), 3)
        with pytest.raises(_host.HostError, match="no CUDA device"):
            _host.Generator(d)
    with pytest.raises(_host.HostError, match="config.json"):
        _host.Generator("/nonexistent/model/dir")


# ---- continuous batching (host/scheduler.*) over the deterministic fake engine ---------------------------------
def _fake_sequential(prompt, max_new, vocab, eos):
    out, last, pos = [], prompt[-1], len(prompt) - 1
    while True:
        nxt = (31 * last + 7 * pos + 3) % vocab
        if nxt == eos:
            return out, "stop"
        out.append(nxt)
        if len(out) >= max_new:
            return out, "length"
        last, pos = nxt, pos + 1


def test_scheduler_batches_ragged_requests_and_matches_sequential_generation():
    from gabby_b200 import _host
    vocab, eos = 997, 13
    rng = np.random.default_rng(3)
    reqs = [(rng.integers(1, vocab, size=int(rng.integers(1, 60))).tolist(), int(rng.integers(1, 40))) for _ in range(23)]
    s = _host.Scheduler(None, eos_ids=[eos], max_batch=4, max_positions=128, max_prefill_tokens=128, num_pages=64, page_size=16,
                        fake_vocab=vocab)
    ids = [s.submit(p, m) for p, m in reqs[:10]]
    for _ in range(5):                      # requests keep arriving while others are running
        s.step()
    ids += [s.submit(p, m) for p, m in reqs[10:]]
    s.drain()
    for rid, (p, m) in zip(ids, reqs):
        toks, fin, done = s.result(rid)
        exp, exp_fin = _fake_sequential(p, m, vocab, eos)
        assert done and toks == exp and fin == exp_fin, (rid, fin, exp_fin)
    st = s.stats()
    assert st["max_concurrent"] == 4                      # the batch filled up
    assert st["decode_calls"] < sum(len(_fake_sequential(p, m, vocab, eos)[0]) for p, m in reqs)   # steps were shared
    assert st["free_pages"] == 64                         # every page came back
    assert st["prefill_tokens"] >= sum(len(p) for p, _ in reqs)
    s.close()


def test_scheduler_preempts_and_recomputes_when_the_kv_pool_runs_out():
    from gabby_b200 import _host
    vocab, eos = 1009, 1008   # this eos is rarely produced: sequences run to their length limit
    # 6 pages of 16 tokens; three sequences of 20 + 40 tokens each need 4 pages apiece -> cannot all grow at once
    s = _host.Scheduler(None, eos_ids=[eos], max_batch=3, max_positions=64, max_prefill_tokens=64, num_pages=6, page_size=16,
                        fake_vocab=vocab)
    prompts = [[(7 * i + j) % vocab + 1 for j in range(20)] for i in range(3)]
    ids = [s.submit(p, 40) for p in prompts]
    s.drain()
    for rid, p in zip(ids, prompts):
        toks, fin, done = s.result(rid)
        exp, exp_fin = _fake_sequential(p, 40, vocab, eos)
        assert done and toks == exp and fin == exp_fin
    st = s.stats()
    assert st["preemptions"] >= 1 and st["free_pages"] == 6
    s.close()


def test_scheduler_rejects_requests_that_can_never_run():
    from gabby_b200 import _host
    s = _host.Scheduler(None, eos_ids=[5], max_batch=2, max_positions=32, max_prefill_tokens=32, num_pages=4, page_size=16, fake_vocab=100)
    with pytest.raises(_host.HostError):
        s.submit([], 4)
    with pytest.raises(_host.HostError):
        s.submit([1] * 30, 8)          # prompt + max_new > max_positions
    with pytest.raises(_host.HostError):
        s.submit([1, 2, 3], 0)
    assert s.step() == 0               # idle
    s.close()

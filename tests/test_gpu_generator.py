"""GPU tests of the drop-in boundary: gabby::inference::Llama3Generator (gabby_b200/host/generator.*)
loaded from an HF-layout directory exactly as gabby's InferenceService does
(/root/reference/src/service.cc:120-124: Llama3Generator::Load(LoadConfig(model_dir))), checked against
the CPU oracle on the same directory."""
import json
import os
import tempfile

import numpy as np
import pytest

from gabby_b200 import synth

pytestmark = pytest.mark.gpu


def _oracle_ids(d, arch, prompt, n):
    from oracle import pyoracle as po
    m = po.OracleModel.from_dir(d, arch, 256)
    ids, margins = m.seq(po.ORC_KV_BF16).greedy(prompt, n)
    return ids.tolist(), float(margins.min())


@pytest.mark.parametrize("preset,shards", [("tiny", 1), ("tiny128", 3)])
def test_generator_token_path_matches_oracle(preset, shards):
    from gabby_b200 import _host
    arch = synth.preset(preset)
    with tempfile.TemporaryDirectory() as d:
        synth.write_model_dir(d, arch, 41, shards=shards)
        if shards > 1:   # the oracle helper reads a single file: write the same weights unsharded next to it
            d1 = os.path.join(d, "single")
            synth.write_model_dir(d1, arch, 41)
        else:
            d1 = d
        prompt = synth.synth_prompt(17, arch.vocab_size, arch.bos_token_id, 6)
        want, margin = _oracle_ids(d1, arch, prompt, 24)
        gen = _host.Generator(d, max_positions=256, max_new_tokens=24)
        for device_loop in (False, True):
            ids, fin = gen.generate_ids(prompt, 24, device_loop=device_loop)
            assert ids.tolist() == want, (device_loop, margin)
            assert fin == "length"
        gen.close()


def test_generator_stops_on_eos_and_reports_finish_reason():
    from gabby_b200 import _host
    arch = synth.preset("tiny")
    with tempfile.TemporaryDirectory() as d:
        synth.write_model_dir(d, arch, 41)
        prompt = synth.synth_prompt(17, arch.vocab_size, arch.bos_token_id, 6)
        want, _ = _oracle_ids(d, arch, prompt, 12)
        stop_at = 5
        eos = want[stop_at]
        first = want.index(eos)
        with open(os.path.join(d, "generation_config.json"), "w") as f:
            json.dump({"eos_token_id": [eos, 999]}, f)
        gen = _host.Generator(d, max_positions=256, max_new_tokens=12)
        for device_loop in (False, True):
            ids, fin = gen.generate_ids(prompt, 12, device_loop=device_loop)
            assert fin == "stop" and ids.tolist() == want[:first]
        gen.close()


def test_generate_request_to_message_and_errors():
    from gabby_b200 import _host
    arch = synth.preset("tiny")
    with tempfile.TemporaryDirectory() as d:
        synth.write_model_dir(d, arch, 41)
        gen = _host.Generator(d, max_positions=128, max_new_tokens=8)
        a = gen.generate("You are a helpful assistant.", "Hello!")
        b = gen.generate("You are a helpful assistant.", "Hello!")
        assert isinstance(a, str) and a == b                      # deterministic greedy decoding
        with pytest.raises(_host.HostError, match="does not fit"):
            gen.generate("x" * 200, "y")                          # byte-fallback prompt longer than the context
        with pytest.raises(_host.HostError, match="token id out of range"):
            gen.generate_ids([arch.vocab_size + 5], 4)
        gen.close()
    with tempfile.TemporaryDirectory() as d:
        synth.write_model_dir(d, arch, 41)
        os.remove(os.path.join(d, "tokenizer.json"))              # LoadConfig needs all five JSON files (config.cc:13-17)
        with pytest.raises(_host.HostError, match="tokenizer.json"):
            _host.Generator(d)


def test_continuous_batching_scheduler_matches_per_sequence_oracle():
    """host/scheduler.*: requests of different lengths arrive over time, share ragged prefill and decode steps of the real
    engine, finish independently -- every result must equal the CPU oracle's greedy continuation of that prompt alone."""
    from gabby_b200 import _host
    from oracle import pyoracle as po
    from tests.helpers import synth_tensors, make_engine
    arch, tensors = synth_tensors("tiny128", None, 77)
    eng = make_engine(arch, tensors, max_batch=4, max_positions=96, page_size=16, num_pages=20, max_prefill_tokens=96)
    om = po.OracleModel(arch, tensors, 96)
    lens = [5, 33, 17, 1, 48, 9, 26]
    news = [12, 6, 20, 9, 5, 16, 8]
    prompts = [synth.synth_prompt(n, arch.vocab_size, arch.bos_token_id, 400 + i) for i, n in enumerate(lens)]
    expect = [om.seq(po.ORC_KV_BF16).greedy(p, m)[0].tolist() for p, m in zip(prompts, news)]
    stop_id = expect[0][2]                                   # request 0 must stop right before this token
    want0 = expect[0][:expect[0].index(stop_id)]
    sched = _host.Scheduler(eng, eos_ids=[stop_id], max_batch=4, max_positions=96, max_prefill_tokens=96, num_pages=20, page_size=16)
    ids = [sched.submit(p, m) for p, m in zip(prompts[:4], news[:4])]
    for _ in range(3):
        sched.step()
    ids += [sched.submit(p, m) for p, m in zip(prompts[4:], news[4:])]
    sched.drain()
    for i, rid in enumerate(ids):
        toks, fin, done = sched.result(rid)
        exp = expect[i]
        if stop_id in exp:
            exp = exp[:exp.index(stop_id)]
            assert fin == "stop", (i, fin)
        else:
            assert fin == "length", (i, fin)
        assert done and toks == exp, (i, toks, exp)
    assert sched.result(ids[0])[0] == want0
    st = sched.stats()
    assert st["max_concurrent"] >= 3 and st["free_pages"] == 20
    sched.close()
    eng.close()

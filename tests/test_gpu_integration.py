"""The drop-in boundary against the ACTUAL reference: integration/_build/service_b200 is gabby's own InferenceService, HTTP
server, router and JSON code (compiled from /root/reference/src by integration/Makefile) with
`B200Llama3Generator : gabby::inference::Generator` injected through the constructor gabby's tests use
(/root/reference/src/service.cc:126-129). The program replays /root/reference/src/service_test.cc:28-57 -- POST
/v1/chat/completions through gabby's http::PostJson -- and the assistant's content must be the CPU oracle's greedy
continuation of the chat prompt, detokenized."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from gabby_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "integration", "_build", "service_b200")


def test_integration_sources_reach_the_engine_only_through_the_c_abi():
    """CPU: the adapter includes gabby's header and this repo's C header, nothing else of this repo."""
    for name in ("b200_generator.h", "b200_generator.cc", "service_b200_main.cc"):
        text = open(os.path.join(ROOT, "integration", name)).read()
        incs = [line.split('"')[1] for line in text.splitlines() if line.startswith('#include "')]
        for inc in incs:
            assert inc in ("b200_generator.h", "gabby_b200_host.h") or inc.split("/")[0] in ("inference", "http", "json", "utils", "service.h"), inc
    mk = open(os.path.join(ROOT, "integration", "Makefile")).read()
    assert "$(REF)/src/service.cc" in mk and "-lgabby_host -lb2l" in mk


@pytest.mark.gpu
def test_gabby_service_answers_with_the_b200_generator():
    if not os.path.exists(BIN):
        pytest.skip("integration/_build/service_b200 is built in the dev container (needs /root/reference)")
    from gabby_b200 import _host
    from oracle import pyoracle as po
    arch = synth.preset("tiny")
    system = "You are a helpful assistant."
    with tempfile.TemporaryDirectory() as d:
        synth.write_model_dir(d, arch, 41)
        # what the answer must be: chat template + tokenizer of the host layer (checked against HF tokenizers in
        # tests/test_host_layer.py), greedy continuation by the CPU oracle, detokenized
        tok = _host.Tokenizer(open(os.path.join(d, "tokenizer.json")).read())
        om = po.OracleModel.from_dir(d, arch, 256)
        eos = set(arch.eos_token_ids)
        user = want = margins = None
        for cand in ["Hello!", "Hi there", "Tell me a story", "What is 2 + 2?", "Good morning", "Why is the sky blue?", "abc", "Name a colour",
                     "Say something", "One more question", "How are you?", "Thanks"]:
            prompt = tok.chat_prompt(system, cand)
            ids, m = om.seq(po.ORC_KV_BF16).greedy(np.asarray(prompt, dtype=np.int32), 16)
            keep = []
            for t in ids.tolist():
                if t in eos:
                    break
                keep.append(t)
            text = tok.detokenize_bytes(keep)
            # gabby's JSON writer does not escape strings and its reader ends a string at a newline
            # (/root/reference/src/json/json.h:172-173, parser.cc:111-121): pick an answer its own envelope can carry
            if text and all(b >= 0x20 and b not in (0x22, 0x5C) for b in text) and float(m.min()) > 1e-3:
                user, want, margins = cand, text, m
                break
        assert user is not None, "no candidate request has an answer gabby's JSON envelope can carry"
        r = subprocess.run([BIN, d, system, user, "3"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
        lines = [x.split(" ") for x in r.stdout.splitlines() if x.startswith("RESPONSE ")]
        assert len(lines) == 3
        for _, obj, role, *hexed in lines:
            assert obj == "chat.completion" and role == "assistant"
            got = bytes.fromhex(hexed[0] if hexed else "")
            assert got == want, (got, want, float(margins.min()))

"""Generate tests/golden/*.npz from HF transformers (fp32, CPU) on synthetic model dirs.

Run in the dev container:  python tests/golden/make_golden.py
The reference (dhconnelly/gabby) has no forward pass and no golden vectors for this path, so
the oracle is pinned to HF `LlamaForCausalLM` instead. The fixtures hold
ONLY HF outputs; weights are regenerated from the seed by gabby_b200/synth.py.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gabby_b200 import synth  # noqa: E402

CASES = [
    # (fixture name, preset, layers, seed, prompt_len, n_new)
    ("tiny_s1234", "tiny", None, 1234, 24, 16),
    ("tiny128_s77", "tiny128", None, 77, 19, 12),
    ("w1b_l2_s5", "1b", 2, 5, 12, 6),  # full 1B width (H=2048, V=128256), 2 layers
    ("w3b_l2_s34", "3b", 2, 34, 12, 6),  # Llama-3.2-3B width: H=3072, 24 query / 8 kv heads (GQA group 3), hd 128, tied head
    ("w8b_l2_s9", "8b", 2, 9, 12, 6),    # Llama-3.1-8B width: H=4096, I=14336, GQA group 4, untied lm_head, rope factor 8
    ("w70b_l2_s71", "70b", 2, 71, 10, 4),  # Llama-3.1-70B width: H=8192, I=28672, 64 query / 8 kv heads (GQA group 8)
]


def run_case(name, preset, layers, seed, n_prompt, n_new):
    import torch
    from transformers import AutoConfig, LlamaForCausalLM

    os.environ["HF_HUB_OFFLINE"] = "1"
    arch = synth.preset(preset, layers)
    with tempfile.TemporaryDirectory() as d:
        synth.write_model_dir(d, arch, seed)
        model = LlamaForCausalLM.from_pretrained(d, torch_dtype=torch.float32, attn_implementation="eager")
        model.eval()
        prompt = synth.synth_prompt(n_prompt, arch.vocab_size, arch.bos_token_id, seed + 1)
        ids = torch.tensor(prompt[None, :].astype(np.int64))
        with torch.no_grad():
            out = model(ids, output_hidden_states=True, use_cache=True)
            logits = out.logits[0].numpy().astype(np.float32)  # [n, V]
            hs = np.stack([h[0].numpy() for h in out.hidden_states]).astype(np.float32)  # [L+1, n, H]
            # greedy continuation with KV cache
            past = out.past_key_values
            nxt = int(out.logits[0, -1].argmax())
            new_ids, step_logits = [nxt], []
            for _ in range(n_new - 1):
                o = model(torch.tensor([[nxt]]), past_key_values=past, use_cache=True)
                past = o.past_key_values
                step_logits.append(o.logits[0, -1].numpy().astype(np.float32))
                nxt = int(o.logits[0, -1].argmax())
                new_ids.append(nxt)
        inv_freq = model.model.rotary_emb.inv_freq.numpy().astype(np.float32)
    V = arch.vocab_size
    # keep fixtures small: full logits only for small vocabularies, else a strided sample + top-8
    top = np.argsort(-logits, axis=1)[:, :8].astype(np.int32)
    if V > 4096:
        cols = np.unique(np.concatenate([np.arange(0, V, 97), top.reshape(-1)])).astype(np.int32)
        cols = cols[:4096]
    else:
        cols = np.arange(V, dtype=np.int32)
    H = arch.hidden_size
    hcols = np.arange(0, H, max(1, H // 256), dtype=np.int32)
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", name + ".npz"),
        preset=preset, layers=-1 if layers is None else layers, seed=seed,
        prompt=prompt, logit_cols=cols, logits=logits[:, cols], top8=top,
        hidden_cols=hcols, hidden=hs[:, :, hcols], greedy_ids=np.array(new_ids, dtype=np.int32),
        step_logits=np.stack(step_logits)[:, cols] if step_logits else np.zeros((0, cols.size), np.float32),
        inv_freq=inv_freq,
    )
    print(name, "ok: greedy", new_ids)


if __name__ == "__main__":
    only = sys.argv[1:] or None
    for c in CASES:
        if only is None or c[0] in only:
            run_case(*c)

"""One tensor-parallel rank (spawned by tests/test_gpu_tp.py): builds its shard of the engine from the
FULL HF tensors (the engine keeps its slice), runs prefill + a greedy loop, reports ids and logits."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(rank, world, nccl_id, preset, layers, seed, prompt, n_new, max_batch, out_q, transport="peer", n_single=0):
    try:
        os.environ["B2L_TP_TRANSPORT"] = transport
        from gabby_b200 import _capi, _host, synth
        arch = synth.preset(preset, layers)
        eng = _capi.Engine(arch, _host.rope_table(arch, 256), max_batch=max_batch, max_positions=256, max_prefill_tokens=128,
                           device=rank, tp_rank=rank, tp_size=world, nccl_id=nccl_id)
        for name, shape, scale, off in synth.tensor_specs(arch):
            if rank == 0:   # exercise both ways of getting a shard: host upload (rank 0) and on-device generation
                eng.upload(name, synth.gen_tensor_bits(name, int(np.prod(shape)), scale, off, seed), shape)
            else:
                eng.synth(name, shape, synth.tensor_seed(name, seed), scale, off)
        eng.finalize()
        n_seq = len(prompt)
        bt = np.arange(n_seq * eng.max_blocks, dtype=np.int32).reshape(n_seq, eng.max_blocks)
        first = eng.prefill(prompt, [0] * n_seq, bt)
        logits0 = eng.logits(0, n_seq)                       # collective: all ranks call it
        ids, ms = eng.decode_loop(first, [len(p) for p in prompt], bt, n_new)
        logits1 = eng.logits(0, n_seq)
        singles, tok, pos = [], ids[-1].copy(), [len(p) + n_new for p in prompt]
        for _ in range(n_single):            # a few more tokens through the per-step call (host token in, id out)
            tok = eng.decode(tok, pos, bt)
            pos = [x + 1 for x in pos]
            singles.append(int(tok[0]))
        info = eng.info()
        out_q.put((rank, "ok", first.tolist(), ids.tolist(), logits0, logits1, int(info.weight_bytes), int(info.decode_mode), int(info.tp_transport), singles))
        eng.close()
    except Exception as e:  # noqa: BLE001
        out_q.put((rank, "error", repr(e)))

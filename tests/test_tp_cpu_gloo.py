"""CPU tests of the N>1 paths with world_size 2 over gloo (127.0.0.1):
  * the tensor-parallel shard windows the CUDA engine uses (b2l_shard_window, pure host arithmetic):
    every sharded element owned exactly once, and column/row-parallel matmuls recombine with an all-reduce;
  * bench.py's multi-rank conventions for the reference arm (rank 0 prints, others exit 0 silently)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from gabby_b200 import _capi, build, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        arch = synth.preset("tiny128")
        seed = 5
        problems = []
        for name, shape, scale, off in synth.tensor_specs(arch):
            full = synth.bf16_bits_to_f32(synth.gen_tensor_bits(name, int(np.prod(shape)), scale, off, seed)).reshape(
                shape if len(shape) == 2 else (1, shape[0]))
            r0, nr, c0, nc = _capi.shard_window(arch, name, shape, rank, world)
            cover = torch.zeros(full.shape, dtype=torch.int32)
            cover[r0:r0 + nr, c0:c0 + nc] += 1
            dist.all_reduce(cover)
            replicated = ("norm" in name) or name == "model.embed_tokens.weight"
            want = world if replicated else 1
            if not bool((cover == want).all()):
                problems.append((name, "coverage", int(cover.min()), int(cover.max())))
            # the algebra the shards imply: row-sharded outputs concatenate, column-sharded partial sums add up
            x = np.random.default_rng(7).standard_normal(full.shape[1]).astype(np.float32)
            shard = full[r0:r0 + nr, c0:c0 + nc]
            if not replicated:
                if nc == full.shape[1]:      # output (row) sharded: gather
                    y = torch.zeros(full.shape[0], dtype=torch.float32)
                    y[r0:r0 + nr] = torch.from_numpy(shard @ x)
                else:                        # input (column) sharded: partial products, all-reduce(sum)
                    y = torch.from_numpy(shard @ x[c0:c0 + nc])
                dist.all_reduce(y)
                ref = full @ x
                if np.abs(y.numpy() - ref).max() > 1e-4 * max(1.0, np.abs(ref).max()):
                    problems.append((name, "algebra"))
        q.put((rank, problems))
    finally:
        dist.destroy_process_group()


def test_tp_shard_windows_partition_and_recombine_world2_gloo():
    import torch.multiprocessing as tmp
    from gabby_b200 import build
    build.build_cuda()
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, problems in res:
        assert problems == [], (rank, problems)


def test_shard_window_values_for_llama31_8b_tp8():
    from gabby_b200 import _capi, synth
    a = synth.preset("8b")
    assert _capi.shard_window(a, "model.layers.3.self_attn.q_proj.weight", (4096, 4096), 5, 8) == (5 * 512, 512, 0, 4096)
    assert _capi.shard_window(a, "model.layers.3.self_attn.k_proj.weight", (1024, 4096), 5, 8) == (5 * 128, 128, 0, 4096)
    assert _capi.shard_window(a, "model.layers.3.self_attn.o_proj.weight", (4096, 4096), 5, 8) == (0, 4096, 5 * 512, 512)
    assert _capi.shard_window(a, "model.layers.3.mlp.down_proj.weight", (4096, 14336), 7, 8) == (0, 4096, 7 * 1792, 1792)
    assert _capi.shard_window(a, "lm_head.weight", (128256, 4096), 1, 8) == (16032, 16032, 0, 4096)
    assert _capi.shard_window(a, "model.norm.weight", (4096,), 3, 8) == (0, 1, 0, 4096)
    with pytest.raises(_capi.B2lError, match="divide by tp_size"):
        _capi.shard_window(a, "model.norm.weight", (4096,), 0, 3)


def test_bench_reference_arm_multi_rank_convention():
    """Under torchrun (N > 1) rank 0 alone runs and prints the reference line; other ranks exit 0 silently."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""

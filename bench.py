#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: Llama-3.2-1B bf16 batch-1 decode tok/s (+ % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one greedy decode token of one sequence at context >= 512 (configs[1]). N > 1 (under
torchrun) runs N independent replicas -- the 1B path does not shard (SURVEY.md section 8e:
"replicas only"); `value` is the aggregate over ranks, scaling "weak".

  value    device-resident greedy loop (b2l_decode_loop: token feedback stays in HBM), CUDA-event
           time on the engine's stream, max over ranks
  e2e      the same steps through the per-step C-ABI call b2l_decode with HOST buffers: H2D of
           token/position/block table and D2H of the next id inside the timed region
  roofline algorithmic bytes per token / step time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference: the scalar-structured fp32 C++ oracle (oracle/, a port -- the
           reference has no forward pass) on the host cores, bounded sample

Weights are synthetic (gabby_b200/synth.py counter hash), generated on-device for the CUDA arm
and by the oracle's C generator for the CPU arm -- bit-identical (tests/test_gpu_parity.py).

Besides the headline the same JSON line carries, measured in the same invocation:
  tp            BASELINE configs[3]: Llama-3.1-8B bf16 tensor-parallel over the N ranks (strong scaling: the model is
                fixed), batch 1 and batch 32 decode from a 4096-token context; with N > 1 rank 0 also times TP=1 so that
                `eff_vs_tp1` compares numbers of one run on one box
  secondary     (N = 1 only) BASELINE configs[2]: Llama-3.2-3B batch 8, 2048-token prefill + decode
  parity_check  greedy ids of two-layer same-width variants of every model above, run through the same CUDA paths
                right here, against oracle ids committed in tests/golden/bench_parity.json
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018
CTX0 = 512
PAGE = 16


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_decode_sample(arch, n_prefill: int, n_steps: int, warm: int):
    """Oracle (port) decode tok/s on the host cores. Returns (tok/s, cores, description)."""
    from gabby_b200 import synth
    from oracle import pyoracle as po
    # torchrun exports OMP_NUM_THREADS=1; the CPU leg is meant to use every host core this process may run on
    try:
        usable = len(os.sched_getaffinity(0))
    except AttributeError:
        usable = os.cpu_count() or 1
    po.lib().orc_set_num_threads(usable)
    tensors = {n: po.synth_tensor(synth.tensor_seed(n, SEED), int(np.prod(s)), sc, off) for n, s, sc, off in synth.tensor_specs(arch)}
    om = po.OracleModel(arch, tensors, n_prefill + warm + n_steps + 1)
    s = om.seq(po.ORC_KV_BF16)
    prompt = synth.synth_prompt(n_prefill, arch.vocab_size, arch.bos_token_id, SEED + 1)
    lg, _ = s.forward(prompt)
    tok = int(lg[0].argmax())
    for _ in range(warm):
        lg, _ = s.forward([tok]); tok = int(lg[0].argmax())
    t0 = time.perf_counter()
    for _ in range(n_steps):
        lg, _ = s.forward([tok]); tok = int(lg[0].argmax())
    dt = time.perf_counter() - t0
    cores = po.lib().orc_num_threads()
    return n_steps / dt, cores, (f"{n_steps} greedy decode steps of the same 1B weights at context {n_prefill + warm}.."
                                 f"{n_prefill + warm + n_steps} (after a {n_prefill}-token prefill of bench.py's own prompt)")


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from gabby_b200 import synth
    arch = synth.preset("1b")
    steps = max(1, min(args.steps, 64))
    toks, cores, sample = cpu_decode_sample(arch, CTX0, steps, max(1, min(args.warmup, 4)))
    line = {
        "impl": "reference", "metric": "decode_tokens_per_s", "value": toks, "unit": "tok/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 4), "ms_per_step": 1000.0 / toks, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16 weights / fp32 math", "data": "synthetic",
        "config": {"workload": "llama-3.2-1b bf16 batch-1 greedy decode at 512-token context (BASELINE configs[1])", "batch": 1,
                   "context": [CTX0, CTX0 + steps], "note": "CPU: bounded sample of the same workload"},
        "cpu_baseline": {"value": toks, "unit": "tok/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "dhconnelly/gabby has no forward pass (generator.cc:33-38 is a stub); this is the "
                                 "fp32 C++ restatement in oracle/, OpenMP over rows, AVX2 inner loops"},
        "e2e": {"value": toks, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def build_engine(arch, device: int, max_positions: int):
    from gabby_b200 import _capi, _host, synth
    eng = _capi.Engine(arch, _host.rope_table(arch, max_positions), max_batch=1, max_positions=max_positions,
                       page_size=PAGE, max_prefill_tokens=CTX0 + 8, device=device)
    for name, shape, scale, off in synth.tensor_specs(arch):
        eng.synth(name, shape, synth.tensor_seed(name, SEED), scale, off)
    eng.finalize()
    return eng


def share_nccl_id(torch, dist, rank, world, local):
    """128-byte NCCL unique id made by rank 0 and broadcast over torch.distributed (None when world == 1)."""
    if world == 1:
        return None
    from gabby_b200 import _capi
    idt = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(_capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    return bytes(idt.cpu().numpy().tobytes())


def max_over_ranks(torch, dist, world, local, x: float) -> float:
    if world == 1:
        return float(x)
    t = torch.tensor([x], device=f"cuda:{local}", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def measure_tp(torch, dist, preset: str, rank: int, world: int, local: int, batches, ctx0: int, K: int, W: int, with_e2e: bool):
    """Tensor-parallel decode of `preset` over `world` ranks (world 1: one GPU, no collective): per batch size the
    device-timed step (max over ranks), per-GPU HBM fraction, kernels per step. Collective: every rank calls it."""
    from gabby_b200 import _capi, _host, synth
    arch = synth.preset(preset)
    B_max = max(batches)
    max_positions = ctx0 + 2 * (K + W) + 64
    nccl_id = share_nccl_id(torch, dist, rank, world, local)
    eng = _capi.Engine(arch, _host.rope_table(arch, max_positions), max_batch=B_max, max_positions=max_positions, page_size=PAGE,
                       max_prefill_tokens=64, device=local, tp_rank=rank, tp_size=world, nccl_id=nccl_id)
    for name, shape, scale, off in synth.tensor_specs(arch):
        eng.synth(name, shape, synth.tensor_seed(name, SEED), scale, off)
    eng.finalize()
    info = eng.info()
    peak, peak_src = measured_peaks()
    kv_per_tok = 2 * arch.num_hidden_layers * (arch.num_key_value_heads // world) * arch.head_dim * 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {"model": {"8b": "llama-3.1-8b", "70b": "llama-3.1-70b"}.get(preset, preset), "tp": world, "scaling": "strong", "context": ctx0,
           "steps": K, "warmup": W, "kv": f"paged bf16 (zeros), page {PAGE}",
           "transport": {0: "none (single GPU)", 1: "ncclAllReduce fp32 sum after O-proj and down-proj",
                         2: "row-parallel partial sums stored into the peers' NVLink-mapped slabs by the projection epilogue (no all-reduce call)"}[info.tp_transport],
           "decode_mode": int(info.decode_mode), "peak_GBps": peak, "peak_source": peak_src, "by_batch": {}}
    for B in batches:
        bt = np.arange(B * eng.max_blocks, dtype=np.int32).reshape(B, eng.max_blocks)
        tok = synth.synth_prompt(B + 1, arch.vocab_size, arch.bos_token_id, SEED + 2)[1:]
        pos = [ctx0] * B
        eng.decode_loop(tok, pos, bt, W)
        barrier()
        l0 = eng.info().kernels_launched
        _, dev_ms = eng.decode_loop(tok, pos, bt, K)
        launches = eng.info().kernels_launched - l0
        barrier()
        ms_per_step = max_over_ranks(torch, dist, world, local, dev_ms) / K
        bytes_per_step = info.stream_bytes_per_token + B * ((ctx0 + K / 2.0) * kv_per_tok + kv_per_tok)   # per GPU
        achieved = bytes_per_step / (ms_per_step * 1e-3) / 1e9
        rec = {"tok_per_s": B * 1000.0 / ms_per_step, "ms_per_step": ms_per_step, "per_gpu_GBps": achieved, "frac": achieved / peak,
               "hbm_bytes_per_step_per_gpu": bytes_per_step, "kernels_per_step": launches // K,
               # row-parallel exchange: every rank sends its [B][H] fp32 partial (8-byte words) to every peer after O-proj and down-proj
               "nvlink_bytes_per_step_per_gpu": (2 * arch.num_hidden_layers * B * arch.hidden_size * 8 * (world - 1)) if world > 1 else 0}
        if with_e2e:
            cur, p = tok.copy(), list(pos)
            for _ in range(min(W, 4)):
                cur = eng.decode(cur, p, bt); p = [x + 1 for x in p]
            barrier()
            cur, p = tok.copy(), list(pos)
            t0 = time.perf_counter()
            for _ in range(K):
                cur = eng.decode(cur, p, bt); p = [x + 1 for x in p]
            e2e_s = max_over_ranks(torch, dist, world, local, time.perf_counter() - t0)
            rec["e2e_tok_per_s"] = B * K / e2e_s
            rec["h2d_bytes_per_step"] = B * (12 + 4 * eng.max_blocks)
            rec["d2h_bytes_per_step"] = 4 * B
        out["by_batch"][str(B)] = rec
    eng.close()
    return out


def parity_leg(torch, dist, leg: str, rank: int, world: int, local: int):
    """Run the two-layer same-width variant named by `leg` (tests/golden/bench_parity.json) through the CUDA path -- tensor
    parallel over `world` ranks when world > 1 -- and compare its greedy ids with the committed oracle ids."""
    from gabby_b200 import _capi, _host, synth
    with open(os.path.join(ROOT, "tests", "golden", "bench_parity.json")) as f:
        gold = json.load(f)
    g = gold["legs"][leg]
    arch = synth.preset(g["preset"], g["layers"])
    lens, n_seq = g["prompt_lens"], len(g["prompt_lens"])
    n_new = len(g["ids"][0]) - 1
    max_positions = max(lens) + n_new + 16
    nccl_id = share_nccl_id(torch, dist, rank, world, local)
    eng = _capi.Engine(arch, _host.rope_table(arch, max_positions), max_batch=n_seq, max_positions=max_positions, page_size=PAGE,
                       max_prefill_tokens=sum(lens) + 8, device=local, tp_rank=rank, tp_size=world, nccl_id=nccl_id)
    for name, shape, scale, off in synth.tensor_specs(arch):
        eng.synth(name, shape, synth.tensor_seed(name, gold["seed"]), scale, off)
    eng.finalize()
    exact = g["prefill"] == "exact"
    eng.set_prefill_mode(0 if exact else 1)
    prompts = [synth.synth_prompt(n, arch.vocab_size, arch.bos_token_id, sd) for n, sd in zip(lens, g["prompt_seeds"])]
    bt = np.arange(n_seq * eng.max_blocks, dtype=np.int32).reshape(n_seq, eng.max_blocks)
    first = eng.prefill(prompts, [0] * n_seq, bt)
    ids, _ = eng.decode_loop(first, lens, bt, n_new)
    mode = int(eng.info().decode_mode)
    eng.close()
    tie = 0.0 if exact else 6e-2      # bf16 activations in the prefill: logits agree to ~3e-2, a smaller margin may flip
    matched, total, status, detail = 0, 0, "ok", []
    for i in range(n_seq):
        got = [int(first[i])] + [int(x) for x in ids[:, i]]
        want, margins = g["ids"][i], g["margins"][i]
        total += len(want)
        n_ok = next((j for j, (a, b) in enumerate(zip(got, want)) if a != b), len(want))
        matched += n_ok
        if n_ok < len(want):
            if margins[n_ok] < tie:
                status = "ok_tie" if status == "ok" else status
                detail.append(f"seq {i}: diverged at step {n_ok} where the oracle margin is {margins[n_ok]:.3g} (< {tie})")
            else:
                status = "FAIL"
                detail.append(f"seq {i}: id {n_ok} is {got[n_ok]}, oracle {want[n_ok]}, margin {margins[n_ok]:.3g}")
    return {"status": status, "ids_matched": f"{matched}/{total}", "prefill": g["prefill"], "decode_mode": mode, "tp": world, "detail": detail}


def run_8b_tp(args):
    """`--workload 8b-tp | 70b-tp`: BASELINE configs[3] / configs[4] as the headline of the line (the default run carries
    the 8B numbers in its `tp` record)."""
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    big = args.workload == "70b-tp"
    K, W, B, ctx0 = args.steps, max(3, args.warmup), args.batch, args.context
    with ClockSampler(local) as clk:
        tp = measure_tp(torch, dist, "70b" if big else "8b", rank, world, local, [B], ctx0, K, W, True)
    rec = tp["by_batch"][str(B)]
    if rank == 0:
        line = {
            "metric": "decode_tokens_per_s", "value": rec["tok_per_s"], "unit": "tok/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"{tp['model']} bf16 tensor-parallel decode, batch {B}, context {ctx0} (BASELINE configs[{4 if big else 3}])",
                       "batch": B, "context": [ctx0, ctx0 + K], "parallelism": f"tp{world}", "kv": tp["kv"], "collective": tp["transport"],
                       "l2": "inputs larger than L2", "decode_mode": tp["decode_mode"]},
            "roofline": {"bound": "hbm", "achieved": rec["per_gpu_GBps"], "peak": tp["peak_GBps"], "unit": "GB/s", "frac": rec["frac"], "traffic": None,
                         "peak_source": tp["peak_source"], "bytes_per_step_per_gpu": rec["hbm_bytes_per_step_per_gpu"],
                         "kernel": f"decode step = {rec['kernels_per_step']} kernel launches per step; per-GPU fraction"},
            "cpu_baseline": None,
            "e2e": {"value": rec["e2e_tok_per_s"], "unit": "tok/s", "h2d_bytes_per_step": rec["h2d_bytes_per_step"], "d2h_bytes_per_step": rec["d2h_bytes_per_step"]},
            "gpu_launches": int(rec["kernels_per_step"]) * K, "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_3b_b8(local: int, K: int, W: int):
    """BASELINE configs[2]: Llama-3.2-3B bf16, batch 8, 2048-token prefill (tcgen05 GEMMs + flash attention) and K greedy
    decode steps of the whole batch. One GPU."""
    import torch
    from gabby_b200 import _capi, _host, synth
    B, S = 8, 2048
    arch = synth.preset("3b")
    max_positions = S + 2 * (K + W) + 64
    eng = _capi.Engine(arch, _host.rope_table(arch, max_positions), max_batch=B, max_positions=max_positions, page_size=PAGE,
                       max_prefill_tokens=B * S, device=local)
    for name, shape, scale, off in synth.tensor_specs(arch):
        eng.synth(name, shape, synth.tensor_seed(name, SEED), scale, off)
    eng.finalize()
    info = eng.info()
    bt = np.arange(B * eng.max_blocks, dtype=np.int32).reshape(B, eng.max_blocks)
    prompts = [synth.synth_prompt(S, arch.vocab_size, arch.bos_token_id, SEED + 10 + i) for i in range(B)]
    eng.prefill(prompts, [0] * B, bt)                      # warm-up (allocations, tensor maps)
    torch.cuda.synchronize()
    pf = []
    for _ in range(3):
        t0 = time.perf_counter()
        first = eng.prefill(prompts, [0] * B, bt)         # host tokens in, first ids out: end to end
        pf.append(time.perf_counter() - t0)
    pf_s = min(pf)
    lin = sum(int(np.prod(s)) for n, s, _, _ in synth.tensor_specs(arch) if "proj" in n)
    flops = 2.0 * B * S * lin + 2.0 * 2.0 * (S * S / 2.0) * arch.num_attention_heads * arch.head_dim * arch.num_hidden_layers * B
    try:
        tf_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        tf_peak = 1400.0
    pos = [S] * B
    eng.decode_loop(first, pos, bt, W)
    l0 = eng.info().kernels_launched
    ids, dev_ms = eng.decode_loop(first, pos, bt, K)
    launches = eng.info().kernels_launched - l0
    ms_per_step = dev_ms / K
    cur, p = first.copy(), list(pos)
    t0 = time.perf_counter()
    for _ in range(K):
        cur = eng.decode(cur, p, bt); p = [x + 1 for x in p]
    e2e = B * K / (time.perf_counter() - t0)
    mode = int(eng.info().decode_mode)
    eng.close()
    kv_per_tok = 2 * arch.num_hidden_layers * arch.num_key_value_heads * arch.head_dim * 2
    bytes_per_step = info.stream_bytes_per_token + B * ((S + K / 2.0) * kv_per_tok + kv_per_tok)
    peak, peak_src = measured_peaks()
    achieved = bytes_per_step / (ms_per_step * 1e-3) / 1e9
    return {
        "workload": "llama-3.2-3b bf16 batch 8: 2048-token prefill + greedy decode (BASELINE configs[2])", "batch": B, "context": [S, S + K],
        "steps": K, "warmup": W,
        "decode": {"tok_per_s": B * 1000.0 / ms_per_step, "ms_per_step": ms_per_step, "GBps": achieved, "peak_GBps": peak, "frac": achieved / peak,
                   "hbm_bytes_per_step": bytes_per_step, "kernels_per_step": launches // K, "e2e_tok_per_s": e2e, "decode_mode": 0 if B > 1 else mode,   # the megakernel is batch 1 only
                   "h2d_bytes_per_step": B * (12 + 4 * eng.max_blocks), "d2h_bytes_per_step": 4 * B},
        "prefill": {"tokens": B * S, "seconds": pf_s, "tok_per_s": B * S / pf_s, "algorithmic_tflop": flops / 1e12,
                    "achieved_tflops": flops / pf_s / 1e12, "peak_tflops": tf_peak, "frac": flops / pf_s / 1e12 / tf_peak, "bound": "tensor",
                    "note": "tcgen05 GEMMs + flash attention; wall clock incl. H2D of tokens and D2H of the first ids; peak = MEASURED_PEAKS bf16_tflops_sustained"},
    }


def run_3b_b8(args):
    """`--workload 3b-b8`: BASELINE configs[2] as the headline of the line."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    K, W = args.steps, max(3, args.warmup)
    with ClockSampler(local) as clk:
        r = measure_3b_b8(local, K, W)
    d = r["decode"]
    line = {
        "metric": "decode_tokens_per_s", "value": d["tok_per_s"], "unit": "tok/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": d["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": r["workload"], "batch": r["batch"], "context": r["context"], "kv": f"paged bf16, page {PAGE}",
                   "l2": "inputs larger than L2", "decode_mode": d["decode_mode"]},
        "prefill": r["prefill"],
        "roofline": {"bound": "hbm", "achieved": d["GBps"], "peak": d["peak_GBps"], "unit": "GB/s", "frac": d["frac"], "traffic": None,
                     "bytes_per_step": d["hbm_bytes_per_step"], "kernel": f"batch-8 decode step = {d['kernels_per_step']} kernel launches"},
        "cpu_baseline": None,
        "e2e": {"value": d["e2e_tok_per_s"], "unit": "tok/s", "h2d_bytes_per_step": d["h2d_bytes_per_step"], "d2h_bytes_per_step": d["d2h_bytes_per_step"]},
        "gpu_launches": int(d["kernels_per_step"]) * K, "clocks": clk.summary(),
    }
    print(json.dumps(line), flush=True)


def ncu_traffic_per_token():
    """DRAM bytes per token of the megakernel from the committed `ncu --set full` capture, or (None, why) when the kernel source
    has changed since (profiles/megakernel_traffic.json is stamped with the sha256 of mega_decode.cuh it was captured from)."""
    import hashlib
    try:
        with open(os.path.join(ROOT, "profiles", "megakernel_traffic.json")) as f:
            t = json.load(f)
        with open(os.path.join(ROOT, "gabby_b200", "csrc", "mega_decode.cuh"), "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()
        if sha != t["mega_decode_cuh_sha256"]:
            return None, f"stale: {t['source']} was captured from another build of the kernel (commit {t.get('commit', '?')})"
        return float(t["dram_bytes_per_token"]), t["source"]
    except Exception as e:  # noqa: BLE001
        return None, f"no capture: {e}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip the tp / secondary / parity_check records (kernel tuning runs)")
    ap.add_argument("--decode-mode", type=int, default=None)
    ap.add_argument("--workload", default="1b-decode", choices=["1b-decode", "8b-tp", "70b-tp", "3b-b8"],
                    help="1b-decode: BASELINE configs[1], N replicas (default). 8b-tp: configs[3], Llama-3.1-8B tensor-parallel over N GPUs")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--context", type=int, default=4096)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    if args.workload in ("8b-tp", "70b-tp"):
        return run_8b_tp(args)
    if args.workload == "3b-b8":
        return run_3b_b8(args)
    rank, world, local = dist_env()
    K, W = args.steps, max(3, args.warmup)
    import torch
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    from gabby_b200 import synth
    arch = synth.preset("1b")
    max_positions = CTX0 + 2 * (K + W) + 64
    eng = build_engine(arch, local, max_positions)
    if args.decode_mode is not None:
        eng.set_decode_mode(args.decode_mode)
    info = eng.info()
    bt = np.arange(eng.max_blocks, dtype=np.int32)[None, :]
    prompt = synth.synth_prompt(CTX0, arch.vocab_size, arch.bos_token_id, SEED + 1)
    first = eng.prefill([prompt], [0], bt)

    # ---- device-resident loop: `value` ------------------------------------------------------
    eng.decode_loop(first, [CTX0], bt, W)                      # warm-up (also captures the graph)
    barrier()
    launched0 = eng.info().kernels_launched
    with ClockSampler(local) as clk:
        ids, dev_ms = eng.decode_loop(first, [CTX0], bt, K)
        # keep the sampler alive for at least a few samples on short runs
        reps, extra_ms = 0, 0.0
        while dev_ms + extra_ms < 1500.0 and reps < 64:
            _, m = eng.decode_loop(first, [CTX0], bt, K)
            extra_ms += m; reps += 1
    launches = (eng.info().kernels_launched - launched0) // (1 + reps)
    barrier()
    ms_max = max_over_ranks(torch, dist, world, local, dev_ms)
    ms_per_step = ms_max / K
    value = world * 1000.0 / ms_per_step

    # ---- per-step C-ABI calls with host buffers: `e2e` ---------------------------------------
    tok, pos = first.copy(), CTX0
    for _ in range(W):
        tok = eng.decode(tok, [pos], bt); pos += 1
    barrier()
    tok, pos = first.copy(), CTX0
    t0 = time.perf_counter()
    e2e_ids = []
    for _ in range(K):
        tok = eng.decode(tok, [pos], bt); pos += 1
        e2e_ids.append(int(tok[0]))
    e2e_s = max_over_ranks(torch, dist, world, local, time.perf_counter() - t0)
    e2e_value = world * K / e2e_s
    assert e2e_ids == ids[:, 0].tolist(), "device loop and per-step API disagree"
    mega = eng.info().decode_mode == 1
    eng.close()
    # megakernel single-step path: token + position ride in the kernel-argument upload (8 bytes of payload), the block-table
    # row (4 * max_blocks bytes) is re-sent only when a new page is appended (every PAGE steps); multi-kernel path: token,
    # position, slot and the block-table row every step
    h2d = (8 + 4 * eng.max_blocks / PAGE) if info.decode_mode == 1 else (4 + 4 + 4 + 4 * eng.max_blocks)
    d2h = 4

    # ---- roofline ---------------------------------------------------------------------------
    kv_per_tok = 2 * arch.num_hidden_layers * arch.num_key_value_heads * arch.head_dim * 2
    avg_ctx = CTX0 + (K - 1) / 2.0 + 1
    bytes_per_token = info.stream_bytes_per_token + arch.hidden_size * 2 + avg_ctx * kv_per_tok + kv_per_tok
    peak, peak_src = measured_peaks()
    achieved = bytes_per_token / (ms_per_step * 1e-3) / 1e9
    per_token, traffic_src = ncu_traffic_per_token() if mega else (None, "multi-kernel path: no single dominant kernel")
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": per_token * K if per_token else None, "traffic_source": traffic_src,
            "traffic_unit": "DRAM bytes per launch (ncu), same launch as algorithmic_bytes_per_launch",
            "algorithmic_bytes_per_launch": bytes_per_token * K,
            "peak_source": peak_src, "bytes_per_token": bytes_per_token,
            "kernel": (f"persistent decode megakernel, barrier-free dataflow build (1 cooperative launch = {K} tokens)" if mega else
                       f"decode step = CUDA graph of {launches // K} kernels; fraction is for the whole step"),
            "frac_of_8TBps_nominal": achieved / 8000.0}

    # ---- records that ride along: parity of the L=2 variants, 8B tensor parallel, 3B batch 8 --------------------------------
    parity, tp, secondary = {}, None, None
    if not args.headline_only:
        K2, W2 = max(8, min(K, 64)), 4
        if rank == 0:
            parity["1b_l2"] = parity_leg(torch, None, "1b_l2", 0, 1, local)
        barrier()
        tp = measure_tp(torch, dist, "8b", rank, world, local, [1, 32], 4096, K2, W2, False)
        parity[f"8b_l2_tp{world}"] = parity_leg(torch, dist, "8b_l2", rank, world, local)
        if world > 1:
            # strong-scaling efficiency needs the one-GPU number of the SAME run: rank 0 times TP=1 while the others wait
            if rank == 0:
                tp1 = measure_tp(torch, None, "8b", 0, 1, local, [1, 32], 4096, K2, W2, False)
                tp["tp1_same_run"] = {b: {"tok_per_s": r["tok_per_s"], "ms_per_step": r["ms_per_step"], "kernels_per_step": r["kernels_per_step"]}
                                      for b, r in tp1["by_batch"].items()}
                for b, r in tp["by_batch"].items():
                    r["speedup_vs_tp1"] = r["tok_per_s"] / tp1["by_batch"][b]["tok_per_s"]
                    r["eff_vs_tp1"] = r["speedup_vs_tp1"] / world
            barrier()
        else:
            for r in tp["by_batch"].values():
                r["speedup_vs_tp1"], r["eff_vs_tp1"] = 1.0, 1.0
            secondary = {"3b_b8": measure_3b_b8(local, K2, W2)}
            parity["3b_l2"] = parity_leg(torch, None, "3b_l2", 0, 1, local)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    clocks = clk.summary()
    cpu = None
    if not args.no_cpu_baseline:
        toks, cores, sample = cpu_decode_sample(arch, CTX0, 16, 2)
        cpu = {"value": toks, "unit": "tok/s", "cores": cores, "kind": "port", "sample": sample}
    line = {
        "metric": "decode_tokens_per_s", "value": value, "unit": "tok/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "llama-3.2-1b bf16 batch-1 greedy decode at 512-token context (BASELINE configs[1])",
                   "batch": 1, "context": [CTX0, CTX0 + K], "kv": f"paged bf16, page {PAGE}",
                   "parallelism": "replicas" if world > 1 else "single",
                   "l2": "inputs larger than L2: every step streams 2.47 GB of weights (L2 is 126 MB)",
                   "decode_mode": int(info.decode_mode)},
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "tok/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "clocks": clocks,
        "greedy_ids_head": ids[:8, 0].tolist(),
        "parity_check": ("skipped (--headline-only)" if args.headline_only else
                         "ok" if all(v["status"].startswith("ok") for v in parity.values()) else "FAIL"),
        "parity_legs": parity, "tp": tp, "secondary": secondary,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py - the reference's headline workload on B200: FLAN-T5 greedy batch inference,
512-token prompts -> 128 generated tokens, batch 256 (BASELINE.json configs[1]).

A "step" is one pass of the hot path over one 256-prompt batch (tokenised synthetic prompts,
seeded random FLAN-T5-base weights: no checkpoints or datasets exist offline).

  value      generated tokens/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        the same metric through the reference-facing plug-in call
             HuggingFaceModelPredictor._predict_numpy(host numpy batch) -> DataFrame of strings:
             pinned H2D copy + generate + D2H copy + detokenisation inside the timed region
  roofline   cross-attention decode kernel (85 % of decode bytes): algorithmic bytes / launch
             duration vs the measured HBM copy bandwidth; plus the whole decode loop's figure
  cpu_baseline  the reference's CPU path (HF eager fp32 generate through the same predictor)
             on a bounded sample of the same workload, host cores stated

python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
python bench.py --impl reference ...                     (reference arm: the CPU path only)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "generated tokens/sec (and prompts/sec) FLAN-T5-base 512->128"
UNIT = "tokens/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="flan-t5-base")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--new", type=int, default=128)
    ap.add_argument("--lengths", default="full", choices=["full", "alpaca", "uniform"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"],
                    help="numerics contract: bf16 (headline) or the notebook's literal torch_dtype=float16 (fp32 wo, fp32 residual stream)")
    ap.add_argument("--cpu-sample", type=int, default=8, help="prompts in the CPU-baseline sample")
    ap.add_argument("--cpu-timeout", type=int, default=240, help="seconds allowed for the in-run CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def log(msg: str) -> None:
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def effective_cores() -> int:
    """Host cores this process may actually use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(n, 1)


def workload_name(a):
    # BASELINE.json: configs[1] = FLAN-T5-base, batch 256, 512-in/128-out on one B200 (the config the metric is
    # quoted on); configs[3] = the same shape with FLAN-T5-large (HBM-roofline report); configs[0] is the CPU case
    tag = {"flan-t5-base": "BASELINE configs[1]", "flan-t5-large": "BASELINE configs[3] shape", "flan-t5-small": "configs[0] model at the configs[1] shape"}
    std = a.batch == 256 and a.seq == 512 and a.new == 128 and a.lengths == "full"
    note = tag.get(a.model, "custom") if std else "custom shape"
    return f"{a.model} batch {a.batch} {a.seq}-in/{a.new}-out greedy, lengths={a.lengths} ({note})"


# --------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # samples under load = the upper half (idle samples at the edges pull the median down)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU path (reference arm / baseline)
def run_cpu_path(a, steps: int, warmup: int, sample: int):
    """The reference's own CPU implementation of the path: the predictor plug-in driving
    transformers' T5ForConditionalGeneration.generate (eager, fp32) on the host cores, on a
    bounded sample (`sample` prompts per step) of the same workload. Returns tokens/s."""
    import numpy as np
    import torch

    from anyscale_workshop_nyc_2023_b200.synth import SPECS, synthetic_token_batch
    from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir
    from oracle.hf_anchor import load_hf_model
    from anyscale_workshop_nyc_2023_b200.predictor import HuggingFaceModelPredictor
    from transformers import T5Tokenizer

    # eager generate is ~4k tiny ATen ops per decode step: beyond a few dozen threads the
    # per-op fork/join cost dominates, so the intra-op pool is capped (the count used is reported)
    cores = min(effective_cores(), int(os.environ.get("B200T5_CPU_THREADS", "32")))
    torch.set_num_threads(cores)
    log(f"cpu path: {cores} torch threads (host reports {os.cpu_count()} cpus), {sample} prompts/step")
    spec = SPECS[a.model]
    ckpt = checkpoint_dir(a.model, seed=0)
    model = load_hf_model(ckpt, dtype=torch.float32, device="cpu")
    tok = T5Tokenizer.from_pretrained(str(ckpt))
    pred = HuggingFaceModelPredictor(model, tokenizer=tok)
    ids0, mask0 = synthetic_token_batch(1, 16, spec.vocab_size, seed=1, lengths="full")
    pred._predict_numpy({"input_ids": ids0, "attention_mask": mask0}, max_new_tokens=2)  # lazy-init costs, untimed
    times = []
    for s in range(warmup + steps):
        ids, mask = synthetic_token_batch(sample, a.seq, spec.vocab_size, seed=1000 + s, lengths=a.lengths)
        t0 = time.perf_counter()
        df = pred._predict_numpy({"input_ids": ids, "attention_mask": mask, "labels": ids.copy()},
                                 max_new_tokens=a.new, min_new_tokens=a.new)
        dt = time.perf_counter() - t0
        assert len(df) == sample
        log(f"cpu path step {s}: {dt:.1f} s")
        if s >= warmup:
            times.append(dt)
    total = sum(times)
    toks = steps * sample * a.new
    return {"value": toks / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample} prompts/step x {steps} steps of the same {a.seq}->{a.new} workload, HF transformers "
                      f"eager fp32 generate via the predictor plug-in, torch threads={cores}",
            "prompts_per_s": steps * sample / total, "ms_per_step": 1e3 * total / steps}


def main_reference(a):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    steps = max(a.steps, 1)
    warm = min(a.warmup, 1)  # CPU steps are seconds long; at most one warm-up pass
    base = run_cpu_path(a, steps, warm, a.cpu_sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "sampled_prompts_per_step": a.cpu_sample},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "prompts_per_s": base["prompts_per_s"],
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- B200 arm
def main_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    from anyscale_workshop_nyc_2023_b200.synth import SPECS, synthetic_token_batch
    from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir, make_batch_predictor

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a B200; there is no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = SPECS[a.model]
    B, S, T, K, W = a.batch, a.seq, a.new, a.steps, max(a.warmup, 3)

    # rank 0 of the node writes the synthetic checkpoint once; the others wait for it
    if local == 0:
        ckpt = checkpoint_dir(a.model, seed=0)
    if world > 1:
        dist.barrier()
    ckpt = checkpoint_dir(a.model, seed=0)

    log(f"rank {rank}/{world}: checkpoint at {ckpt}; loading the model")
    bp = make_batch_predictor(ckpt, device_map="auto", torch_dtype=torch.float16 if a.dtype == "fp16" else torch.bfloat16)
    from anyscale_workshop_nyc_2023_b200.rayshim.train import _ScoringWorker

    worker = _ScoringWorker(bp._checkpoint, bp._predictor_cls, {**bp._predictor_kwargs, "use_gpu": True}, False)
    predictor = worker.predictor
    model = predictor.model
    # the bench times the named configuration: one static batch per step, whatever its size (larger batches would
    # otherwise be routed through the slot pool, whose occupancy-dependent work is measured by tools/bench_stream.py)
    model.pool_size = max(model.pool_size, a.batch)

    # every rank owns its own shard of batches (dataset sharded by block index, no collective)
    def host_batch(step):
        return synthetic_token_batch(B, S, spec.vocab_size, seed=7919 * (rank + 1) + step, lengths=a.lengths)

    host = [host_batch(s) for s in range(W + K)]
    dev_batches = [(torch.from_numpy(i).to(dev), torch.from_numpy(m).to(dev)) for i, m in host]
    gen_kw = dict(max_new_tokens=T, min_new_tokens=T)  # fixed-length timing: every row emits exactly T tokens

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM, CUDA events
    for s in range(W):
        t_w = time.perf_counter()
        model.generate(input_ids=dev_batches[s][0], attention_mask=dev_batches[s][1], **gen_kw)
        torch.cuda.synchronize()
        if rank == 0:
            log(f"warm-up step {s}: {1e3 * (time.perf_counter() - t_w):.1f} ms {model.stats()}")
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    enc_ms = dec_ms = 0.0
    dec_bytes = enc_flops = 0.0
    e0.record()
    for s in range(W, W + K):
        out = model.generate(input_ids=dev_batches[s][0], attention_mask=dev_batches[s][1], **gen_kw)
        st = model.stats()
        launches += st["kernel_launches"]
        enc_ms += st["encoder_ms"]
        dec_ms += st["decode_ms"]
        dec_bytes += st["decode_algo_bytes"]
        enc_flops += st["encoder_flops"]
        assert out.shape == (B, T + 1)
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()

    if rank == 0:
        log(f"device-resident: {elapsed_ms / K:.1f} ms/step")
    # ---------------- e2e: host numpy batch -> DataFrame of strings through the plug-in
    for s in range(min(W, 2)):
        predictor._predict_numpy({"input_ids": host[s][0], "attention_mask": host[s][1], "labels": host[s][0]}, **gen_kw)
    barrier()
    t0 = time.perf_counter()
    for s in range(W, W + K):
        df = predictor._predict_numpy({"input_ids": host[s][0], "attention_mask": host[s][1], "labels": host[s][0]}, **gen_kw)
        assert len(df) == B and isinstance(df["generated_output"].iloc[0], str)
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    barrier()

    # ---------------- roofline of the dominant kernel (cross-attention decode), measured live
    ca = model.bench_cross_attention(reps=5)

    t_max = torch.tensor([elapsed_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = float(t_max[0]), float(t_max[1])

    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        pk = json.loads(peaks_file.read_text())
        hbm_peak, tf_peak, peak_src = float(pk["hbm_gbs"]), float(pk.get("bf16_tflops_sustained", 1458.8)), "measured (MEASURED_PEAKS.json)"
    else:
        hbm_peak, tf_peak, peak_src = 6650.0, 1400.0, "fallback (B200_PROFILING.md)"
    # DRAM bytes per launch of the roofline kernel from the committed `ncu --set full` capture: only valid for the
    # configuration that was captured (FLAN-T5-base, B=256, S=512, full-length prompts)
    traffic = None
    tf = ROOT / "profiles" / "cross_attn_traffic.json"
    if tf.exists() and a.model == "flan-t5-base" and (B, S, a.lengths) == (256, 512, "full"):
        traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
    from anyscale_workshop_nyc_2023_b200 import roofline  # SURVEY 8(d)'s byte model (tests/test_roofline_cpu.py)

    kv_gb = roofline.cross_attention_bytes_per_launch(spec, [S] * B) * spec.num_decoder_layers / 1e9
    w_gb = 2.0 * roofline.step_weight_elements(spec) / 1e9

    if rank == 0:
        tokens = world * K * B * T
        value = tokens / (elapsed_ms / 1e3)
        ach = ca["bytes_per_launch"] / (ca["ms_per_launch"] / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": workload_name(a), "global_batch": world * B, "prompts_per_rank_step": B,
                       "parallelism": f"dataset sharded over {world} replica(s), no collective",
                       "l2": f"inputs exceed L2 (cross-KV arena {kv_gb:.1f} GB and {w_gb:.2f} GB of decoder weights "
                             "are streamed every step vs 126 MB L2)",
                       "forced_length": "min_new_tokens == max_new_tokens"},
            "prompts_per_s": world * K * B / (elapsed_ms / 1e3),
            "e2e": {"value": tokens / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": 2 * B * S * 8,
                    "d2h_bytes_per_step": B * (T + 1) * 8, "ms_per_step": e2e_ms / K,
                    "api": "HuggingFaceModelPredictor._predict_numpy (numpy batch -> DataFrame[generated_output])"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "attn_decode_kernel<false> (cross-attention decode)",
                         "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algo_bytes_per_launch": ca["bytes_per_launch"], "ms_per_launch": ca["ms_per_launch"]},
            "decode_loop": {"algo_gbytes_per_step_batch": dec_bytes / K / 1e9, "ms": dec_ms / K,
                            "achieved_gbs": dec_bytes / (dec_ms / 1e3) / 1e9, "frac_of_hbm_peak": dec_bytes / (dec_ms / 1e3) / 1e9 / hbm_peak},
            "encoder": {"tflop_per_batch": enc_flops / K / 1e12, "ms": enc_ms / K,
                        "achieved_tflops": enc_flops / (enc_ms / 1e3) / 1e12, "frac_of_bf16_sustained": enc_flops / (enc_ms / 1e3) / 1e12 / tf_peak},
        }
        if world == 1 and not a.no_cpu_baseline:
            log("timing the CPU baseline (bounded sample, own process)")
            cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                   "--model", a.model, "--batch", str(a.batch), "--seq", str(a.seq), "--new", str(a.new),
                   "--lengths", a.lengths, "--cpu-sample", str(a.cpu_sample)]
            env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
            env["CUDA_VISIBLE_DEVICES"] = ""
            try:
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=a.cpu_timeout, env=env)
                ref = json.loads(out.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = ref["cpu_baseline"]
            except (subprocess.TimeoutExpired, IndexError, ValueError, KeyError) as e:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": effective_cores(), "kind": "port",
                                        "sample": f"not finished within {a.cpu_timeout}s ({type(e).__name__})"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    return main_reference(a) if a.impl == "reference" else main_b200(a)


if __name__ == "__main__":
    sys.exit(main())

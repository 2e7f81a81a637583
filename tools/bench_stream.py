"""Natural-EOS batch inference: static `batch_size`-row batches (what BatchPredictor.predict drives, NB:908-913) against
the slot pool (b200t5_generate_stream). Same prompts, same tokens; only the scheduling differs.

  python tools/bench_stream.py [--model flan-t5-base] [--n 2048] [--pool 256] [--lengths alpaca] [--new 128]

Prints one JSON line. Wall-clock with host buffers on both sides (H2D/D2H inside), after one warm-up of each path."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from anyscale_workshop_nyc_2023_b200.modeling import B200T5ForConditionalGeneration  # noqa: E402
from anyscale_workshop_nyc_2023_b200.synth import SPECS, synthetic_token_batch  # noqa: E402
from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir  # noqa: E402


def static_batches(model, ids, mask, pool, T):
    N = ids.shape[0]
    out = np.zeros((N, T + 1), dtype=np.int64)
    lens = np.zeros(N, dtype=np.int32)
    steps = 0
    for lo in range(0, N, pool):
        o, ln = model.generate_host(ids[lo:lo + pool], mask[lo:lo + pool], max_new_tokens=T)
        out[lo:lo + pool, : o.shape[1]] = o
        lens[lo:lo + pool] = ln
        steps += int(model.stats()["decode_steps"])
    return out, lens, steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="flan-t5-base")
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--pool", type=int, default=512, help="decode slots of the pool")
    ap.add_argument("--batch", type=int, default=256, help="rows per static batch (the notebook's batch_size)")
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--new", type=int, default=128)
    ap.add_argument("--lengths", default="alpaca")
    ap.add_argument("--admit", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--poll", type=int, default=8, help="decode steps between two looks at the finished flags")
    ap.add_argument("--only-pool", action="store_true")
    a = ap.parse_args()
    spec = SPECS[a.model]
    model = B200T5ForConditionalGeneration.from_pretrained(checkpoint_dir(a.model, 0))
    ids, mask = synthetic_token_batch(a.n, a.seq, spec.vocab_size, seed=3, lengths=a.lengths)
    res = {}
    for name in (("pool",) if a.only_pool else ("static", "pool")):
        best = None
        for rep in range(a.reps + 1):  # rep 0 = warm-up (plan, graphs)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if name == "static":
                out, lens, steps = static_batches(model, ids, mask, a.batch, a.new)
            else:
                out, lens = model.generate_stream(ids, mask, pool=a.pool, admit_min=a.admit, max_new_tokens=a.new, poll_interval=a.poll)
                out = np.pad(out, ((0, 0), (0, a.new + 1 - out.shape[1])))
                steps = int(model.stats()["decode_steps"])
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if rep and (best is None or dt < best):
                best = dt
        res[name] = dict(seconds=best, out=out, lens=lens, steps=steps)
    if a.only_pool:
        r = res["pool"]
        print(json.dumps({"admit": a.admit, "poll": a.poll, "seconds": round(r["seconds"], 4), "decode_steps": r["steps"],
                          "tokens_per_s": round(int(r["lens"].sum()) / r["seconds"], 1)}))
        return
    same = bool((res["static"]["out"] == res["pool"]["out"]).all() and (res["static"]["lens"] == res["pool"]["lens"]).all())
    lens = res["static"]["lens"]
    toks = int(lens.sum())
    line = {
        "workload": f"{a.model}, {a.n} prompts, S={a.seq} lengths={a.lengths}, max_new_tokens={a.new}, natural EOS, static batches of {a.batch}, pool of {a.pool} slots",
        "generated_tokens": toks,
        "len_mean": float(lens.mean()), "len_p50": float(np.median(lens)), "len_p90": float(np.percentile(lens, 90)), "len_max": int(lens.max()),
        "identical_tokens": same,
    }
    for name in ("static", "pool"):
        r = res[name]
        line[name] = {"seconds": round(r["seconds"], 4), "tokens_per_s": round(toks / r["seconds"], 1),
                      "prompts_per_s": round(a.n / r["seconds"], 1), "decode_steps": r["steps"]}
    line["speedup"] = round(res["static"]["seconds"] / res["pool"]["seconds"], 3)
    print(json.dumps(line))


if __name__ == "__main__":
    main()

"""Phase timeline of one decode step of the persistent kernel (csrc/decode_mega.cuh).

Run the bench with B200T5_MEGA_PROF=<step>; the library prints the SM-clock stamps CTA 0 took
before and after every grid barrier of that step to stderr. This script turns one such line
(read from the file given as argv[1]) into a per-phase table: work time and barrier wait."""
import sys

NAMES_LAYER = ["G_qkv", "A_self", "G_o", "E_res1", "G_cq", "A_cross", "G_co", "E_res2", "G_wi", "G_ffo", "E_res3"]


def main(path, ghz=1.965, layers=12):
    line = [l for l in open(path) if l.startswith("MEGA_PROF")][-1]
    st = [int(x) for x in line.split(":")[1].split()]
    names = []
    for l in range(layers):
        names += [f"L{l}.{n}" for n in NAMES_LAYER]
    names += ["G_lm", "finalize", "E_norm0"]
    pairs = list(zip(st[0::2], st[1::2]))  # (before barrier, after barrier)
    agg = {}
    prev_after = None
    rows = []
    for i, (b, a) in enumerate(pairs):
        nm = names[i] if i < len(names) else f"phase{i}"
        work = (b - prev_after) if prev_after is not None else 0
        wait = a - b
        prev_after = a
        rows.append((nm, work / ghz / 1e3, wait / ghz / 1e3))
        key = nm.split(".")[-1]
        w = agg.setdefault(key, [0.0, 0.0, 0])
        w[0] += work / ghz / 1e3
        w[1] += wait / ghz / 1e3
        w[2] += 1
    total = (pairs[-1][1] - pairs[0][0]) / ghz / 1e3
    print(f"step total (first barrier to last): {total:.1f} us, {len(pairs)} barriers")
    print(f"{'phase':10s} {'count':>5s} {'work us':>10s} {'barrier us':>11s} {'avg work':>9s} {'avg barr':>9s}")
    for k, (w, bw, n) in agg.items():
        print(f"{k:10s} {n:5d} {w:10.1f} {bw:11.1f} {w / n:9.2f} {bw / n:9.2f}")


if __name__ == "__main__":
    main(sys.argv[1])

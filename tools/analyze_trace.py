"""Summarise a tools/trace_decode.py timeline: per-kernel in-situ durations, concurrency, idle gaps
inside the decode loop."""
import collections
import json
import re
import sys


def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("b200::", "")
    return n[:60]


def main(path):
    ev = json.load(open(path))
    ev = [e for e in ev if e["dur"] > 0]
    ev.sort(key=lambda e: e["start"])
    names = [short(e["name"]) for e in ev]
    # decode region: after decode_init
    try:
        i0 = max(i for i, n in enumerate(names) if "decode_init" in n)
    except ValueError:
        i0 = 0
    dec = ev[i0 + 1:]
    dn = names[i0 + 1:]
    adv = [i for i, n in enumerate(dn) if "advance_step" in n]
    print(f"{path}: {len(ev)} kernel events, {len(adv)} decode steps traced")
    if len(adv) < 3:
        return
    # steady-state steps: between advance k and advance k+1
    s_lo, s_hi = adv[1] + 1, adv[-1] + 1
    steps = len(adv) - 2
    seg = dec[s_lo:s_hi]
    sn = dn[s_lo:s_hi]
    t0 = seg[0]["start"]
    t1 = max(e["start"] + e["dur"] for e in seg)
    span = t1 - t0
    busy_sum = sum(e["dur"] for e in seg)
    # union of busy intervals
    iv = sorted((e["start"], e["start"] + e["dur"]) for e in seg)
    union, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    for s, e in iv[1:]:
        if s > cur_e:
            union += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    union += cur_e - cur_s
    print(f"  steady steps: {steps}, wall {span / steps:.1f} us/step, sum of kernel time {busy_sum / steps:.1f} us/step, "
          f"GPU busy (union) {union / steps:.1f} us/step, idle {(span - union) / steps:.1f} us/step, "
          f"avg concurrency {busy_sum / union:.2f}")
    agg = collections.OrderedDict()
    for n, e in zip(sn, seg):
        a = agg.setdefault(n, [0.0, 0])
        a[0] += e["dur"]
        a[1] += 1
    for n, (d, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"    {100 * d / busy_sum:5.1f}%  {d / steps:8.1f} us/step  x{c / steps:5.1f}  avg {d / c:7.2f} us  {n}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)

"""ncu launch list (gpu__time_duration.sum per launch, --csv) -> markdown summary per kernel,
split into the encoder pass and one steady-state decode step. Per-launch times under ncu are
cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import re
import sys


def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("b200::", "")
    return n[:70]


def table(title, names, vals, out):
    agg = collections.OrderedDict()
    for n, v in zip(names, vals):
        a = agg.setdefault(n, [0.0, 0])
        a[0] += v
        a[1] += 1
    tot = sum(v[0] for v in agg.values())
    out.append(f"\n### {title}: {tot / 1e3:.1f} us over {len(names)} launches\n")
    out.append("| share | total us | launches | avg us | kernel |\n|---:|---:|---:|---:|---|")
    for n, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        out.append(f"| {100 * v / tot:.1f}% | {v / 1e3:.1f} | {c} | {v / c / 1e3:.2f} | `{n}` |")


def main(path, out_path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [short(r["Kernel Name"]) for r in rows]
    vals = [float(r["Metric Value"].replace(",", "")) for r in rows]  # ns
    init = [i for i, n in enumerate(names) if "decode_init" in n]
    adv = [i for i, n in enumerate(names) if "advance_step" in n]
    out = [f"# Kernel launch summary ({path})", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` over the start of `bench.py` "
           "(FLAN-T5-base, B=256, S=512): first encoder pass and one decode step in steady state."]
    if init:
        first_model = next(i for i, n in enumerate(names) if "prep_mask" in n)
        table("encoder pass (first generate call)", names[first_model:init[0]], vals[first_model:init[0]], out)
    if len(adv) >= 4:
        table("one decode step (4th step)", names[adv[2] + 1: adv[3] + 1], vals[adv[2] + 1: adv[3] + 1], out)
    open(out_path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and mean duration.
Usage: python tools/launch_summary.py launches.csv [other.csv]  (two files: side-by-side mean durations)"""
import collections
import csv
import re
import sys


def load(path, one_step=True):
    """one_step: keep only the launches between the last two optimizer kernels (exactly one training step)."""
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if one_step:
        marks = [i for i, r in enumerate(rows) if "adam_kernel" in r["Kernel Name"]]
        # the AAS-VC step ends with several Adam launches (the duration predictor keeps its own clock): a step boundary is the
        # LAST launch of a run of Adam kernels
        ends = [m for j, m in enumerate(marks) if j + 1 == len(marks) or not all(
            "adam_kernel" in rows[q]["Kernel Name"] or "sqnorm" in rows[q]["Kernel Name"] or "step_advance" in rows[q]["Kernel Name"]
            for q in range(m + 1, marks[j + 1]))]
        if len(ends) >= 2:
            rows = rows[ends[-2] + 1:ends[-1] + 1]
    agg = collections.OrderedDict()
    for row in rows:
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        e = agg.setdefault(name, [0, 0.0])
        e[0] += 1
        e[1] += v
    return agg


def main():
    a = load(sys.argv[1])
    b = load(sys.argv[2]) if len(sys.argv) > 2 else None
    tot = sum(v[1] for v in a.values())
    print(f"# {sys.argv[1]}: {sum(v[0] for v in a.values())} launches, {tot:.1f} us")
    for k, v in sorted(a.items(), key=lambda kv: -kv[1][1]):
        line = f"{k[:84]:84s} n={v[0]:4d} us={v[1]:9.1f} {100 * v[1] / tot:5.1f}% mean={v[1] / v[0]:7.2f}"
        if b is not None and k in b:
            line += f" | other n={b[k][0]:4d} mean={b[k][1] / b[k][0]:7.2f}"
        print(line)


if __name__ == "__main__":
    main()

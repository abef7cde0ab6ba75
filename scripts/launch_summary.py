"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total ms, share.
usage: python scripts/launch_summary.py LIST.csv [--md OUT.md] [--title "..."]"""
import argparse
import collections
import csv
import re


def load(path):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("lr::<unnamed>::", "").replace("void ", "")
        name = re.sub(r"<unnamed>::", "", name)[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    return agg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--md")
    ap.add_argument("--title", default="launch list by kernel")
    ap.add_argument("--own", default="k_,lr::", help="comma separated prefixes of the repo's own kernels")
    a = ap.parse_args()
    agg = load(a.csv)
    total = sum(t for _, t in agg.values())
    own_prefix = tuple(a.own.split(","))
    out = [f"# {a.title}", "", f"Source: `{a.csv}` (cold-cache, serialised per-launch times: shares, not absolutes).", "",
           "| kernel | launches | total ms | share | own |", "|---|---|---|---|---|"]
    own = 0.0
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        mine = k.startswith(own_prefix)
        own += t if mine else 0.0
        if t / total >= 0.002:
            out.append(f"| `{k}` | {n} | {t:.2f} | {100 * t / total:.1f} % | {'yes' if mine else 'library'} |")
    out += ["", f"Total {total:.1f} ms over {sum(n for n, _ in agg.values())} launches; own kernels {100 * own / total:.1f} %."]
    text = "\n".join(out) + "\n"
    if a.md:
        open(a.md, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import collections
import csv
import io
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(io.StringIO("".join(lines))))


def main(path, detail=False):
    rows = load(path)
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    per = []
    for r in rows:
        name = r["Kernel Name"]
        t = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        t *= {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6}.get(unit, 1.0)
        k = re.sub(r"\(.*", "", name)
        k = re.sub(r"^void |tq::\(anonymous namespace\)::|\(anonymous namespace\)::", "", k)
        agg[k][0] += 1
        agg[k][1] += t
        tot += t
        per.append((t, k, r.get("Grid Size", ""), r.get("ID", "")))
    print(f"# {path}: {len(rows)} launches, {tot:.1f} us total (ncu: cold cache, serialised; compare shares)")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d}  {k[:110]}")
    if detail:
        print("# per launch, in launch order")
        for t, k, g, i in per:
            print(f"{i:>5} {t:9.1f} us  grid={g:<14} {k[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], detail=len(sys.argv) > 2)

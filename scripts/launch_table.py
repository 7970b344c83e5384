"""ncu launch-list csv (gpu__time_duration.sum per launch) -> markdown table of per-kernel launches, time and share."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rd = csv.reader(lines)
    header = next(rd)
    ik, iv, iu = header.index("Kernel Name"), header.index("Metric Value"), header.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*", "", r[ik])
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"(at::native::|<unnamed>::|\(anonymous namespace\)::)", "", name)
        rows.append((name[:90], us))
    tot = sum(us for _, us in rows)
    agg = OrderedDict()
    for n, us in rows:
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + us)
    print(f"launches: {len(rows)}   sum of kernel durations: {tot / 1e3:.3f} ms\n")
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f} % |")


if __name__ == "__main__":
    main(sys.argv[1])

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel."""
import csv, re, sys
from collections import defaultdict


def load(path):
    rows = []
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ci = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if len(r) != len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r[ci["Kernel Name"]], us))
    return rows


def short(name):
    m = re.search(r"rnnt::(?:\(anonymous namespace\)::|<unnamed>::)?(\w+(?:<[^>(]*>)?)", name)
    if m:
        return "rnnt::" + m.group(1)
    return "(torch) " + name[-48:]


for path, title in zip(sys.argv[1::2], sys.argv[2::2]):
    rows = load(path)
    agg = defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        a = agg[short(n)]; a[0] += 1; a[1] += us
    tot = sum(us for _, us in rows)
    print(f"== {title}: {len(rows)} launches, {tot / 1e3:.2f} ms of kernel time")
    print("launches    total us   share  kernel")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"{n:8d} {us:11.1f} {100 * us / tot:6.1f}%  {k}")
    print()

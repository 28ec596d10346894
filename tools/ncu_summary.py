"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and, with --list, each launch."""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    gi = H.index("Grid Size") if "Grid Size" in H else None
    out = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        t = float(r[vi].replace(",", ""))
        t = t / 1e6 if r[ui] in ("ns", "nsecond") else (t / 1e3 if r[ui] in ("us", "usecond") else t)
        out.append((r[ki], t, r[gi] if gi is not None else ""))
    return out


def short(name):
    n = name.split("(")[0]
    return n.replace("void ", "").replace("fcmf::", "")[:90]


if __name__ == "__main__":
    data = load(sys.argv[1])
    if "--list" in sys.argv:
        for i, (k, t, g) in enumerate(data):
            print(f"{i:4d} {t:9.4f} ms  {g:>18s}  {short(k)}")
    agg = collections.OrderedDict()
    tot = sum(t for _, t, _ in data)
    for k, t, _ in data:
        a = agg.setdefault(short(k), [0, 0.0])
        a[0] += 1
        a[1] += t
    print(f"# {len(data)} launches, {tot:.3f} ms total (cold-cache, serialised: compare shares)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:9.3f} ms {100 * t / tot:5.1f}%  x{n:3d}  {k}")

"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file F): python tools/ncu_launch_summary.py F "title" """
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
tot = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    us = float(r[14].replace(",", "")) / (1e3 if r[13] in ("ns", "nsecond") else 1.0)
    n, t = tot.get(name, (0, 0.0)); tot[name] = (n + 1, t + us)
total = sum(t for _, t in tot.values())
print(f"{len(rows)} launches of `{sys.argv[2] if len(sys.argv) > 2 else '?'}` under ncu (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares), total {total:.1f} us")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:72]:72s} n={n:4d} us={t:9.1f} share={100 * t / total:5.1f}% avg={t / n:7.1f}")

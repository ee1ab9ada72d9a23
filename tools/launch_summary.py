"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by (kernel, grid): total, count,
average and share.  Usage: python tools/launch_summary.py gpurun_out/launches.csv [top_n]"""
import collections
import csv
import re
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        key = (name, r[ix["Grid Size"]] if "Grid Size" in ix else "")
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised: compare shares)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{a[1]:10.1f} us {a[0]:4d}x  avg {a[1] / a[0]:8.1f}  {100 * a[1] / tot:5.1f}%  {k[0]} {k[1]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)

"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel table (markdown).
Usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<round>_launches.md"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[h], rows[h + 1:]
    ki, vi, ui, gi, bi = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "grid": r[gi], "block": r[bi]})
        a["n"] += 1
        a["us"] += v
    tot = sum(a["us"] for a in agg.values())
    print(f"| kernel | launches | total us | avg us | share | grid (last) | block |\n|---|---:|---:|---:|---:|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
        print(f"| `{k}` | {a['n']} | {a['us']:.1f} | {a['us']/a['n']:.1f} | {a['us']/tot:.3f} | {a['grid']} | {a['block']} |")
    print(f"\ntotal {tot:.1f} us over {sum(a['n'] for a in agg.values())} launches (cold-cache, serialised: compare shares, not absolutes)")


if __name__ == "__main__":
    main(sys.argv[1])

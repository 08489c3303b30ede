"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (stdout, markdown table)."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = r["Kernel Name"]
        for pre in ("tfgnn::", "void "):
            k = k.replace(pre, "")
        k = k.split("(")[0][:70]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | avg us | total ms | share |\n|---|---:|---:|---:|---:|")
    for k, v in agg.items():
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / 1e3:.3f} | {100 * sum(v) / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])

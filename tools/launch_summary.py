"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launches per kernel, and
(with --seq A:B) the launch sequence between two positions.
    python tools/launch_summary.py gpurun_out/launches.csv [--seq 100:200]"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i
            break
    hi = {k: j for j, k in enumerate(h)}
    seq = []
    for r in rows[start + 1:]:
        if len(r) < len(h) or r[hi["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[hi["Metric Value"]].replace(",", ""))
        u = r[hi["Metric Unit"]]
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}[u]
        seq.append((r[hi["Kernel Name"]], v, r[hi["Grid Size"]], r[hi["Block Size"]]))
    return seq


def main():
    seq = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v, g, b in seq:
        agg[n[:110]][0] += 1
        agg[n[:110]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{len(seq)} launches, {tot / 1e3:.2f} ms of kernel time (serialised, cold cache: shares, not absolutes)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{v[1] / 1e3:9.2f} ms {100 * v[1] / tot:5.1f} % {v[0]:5d}  {k}")
    if "--seq" in sys.argv:
        a, b = sys.argv[sys.argv.index("--seq") + 1].split(":")
        for n, v, g, bl in seq[int(a):int(b)]:
            print(f"{v:9.1f} us  {g:16s} {bl:12s} {n[:90]}")


if __name__ == "__main__":
    main()

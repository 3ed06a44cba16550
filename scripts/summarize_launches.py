"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
    python scripts/summarize_launches.py gpurun_out/launches.csv --slim profiles/rNN_launches.csv

--slim writes the launch list itself with only (id, kernel, grid, block, ns) columns."""
import collections
import csv
import re
import sys


def read(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[u]
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = name.replace("cgat::<unnamed>::", "cgat::").replace("void ", "")
        yield row["ID"], name, row["Grid Size"], row["Block Size"], ns


def main():
    path = sys.argv[1]
    rows = list(read(path))
    if "--slim" in sys.argv:
        out = sys.argv[sys.argv.index("--slim") + 1]
        with open(out, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["id", "kernel", "grid", "block", "gpu_time_ns"])
            for r in rows:
                w.writerow([r[0], r[1][:100], r[2], r[3], int(r[4])])
        return
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, name, _, _, ns in rows:
        agg[name[:100]][0] += 1
        agg[name[:100]][1] += ns
    tot = sum(a[1] for a in agg.values())
    own = sum(a[1] for k, a in agg.items() if k.startswith("cgat::"))
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms GPU time (cold-cache, serialised under ncu); "
          f"cgat_b200 kernels {100 * own / tot:.1f}% of it")
    print(f"# {'ms':>9} {'share':>6} {'n':>6} {'avg us':>9}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e6:11.3f} {100 * t / tot:5.1f}% {n:6d} {t / n / 1e3:9.1f}  {k}")


if __name__ == "__main__":
    main()

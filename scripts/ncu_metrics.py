"""Per-kernel summary of an `ncu --set full` capture exported with `--page raw --csv`.

    python scripts/ncu_metrics.py gpurun_out/full_raw.csv profiles/r01_kernel_metrics.json

Writes {kernel: {launches, time_us, dram_read_bytes, dram_write_bytes, traffic_bytes (read+write per launch),
tensor_active_pct, lts_pct, dram_pct, issue_active_pct, registers}} averaged over the captured launches; bench.py
takes `roofline.traffic` from this file."""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3,
        "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "%": 1, "register/thread": 1}
COLS = {
    "time_us": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "tensor_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "registers": "launch__registers_per_thread",
}


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    name = name.replace("cgat::<unnamed>::", "").replace("unnamed>::", "").replace("<unnamed>::", "")
    return name.strip()


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    acc = collections.defaultdict(lambda: collections.defaultdict(float))
    for r in data:
        k = short(r[idx["Kernel Name"]])
        acc[k]["launches"] += 1
        for key, col in COLS.items():
            if col in idx and r[idx[col]] not in ("", "n/a"):
                acc[k][key] += float(r[idx[col]].replace(",", "")) * UNIT.get(units[idx[col]], 1)
    out = {}
    for k, a in acc.items():
        n = a["launches"]
        o = {"launches": int(n)}
        for key in COLS:
            o[key] = round(a[key] / n, 3)
        o["traffic_bytes"] = round((a["dram_read_bytes"] + a["dram_write_bytes"]) / n, 1)
        out[k] = o
    json.dump(out, open(sys.argv[2], "w"), indent=1, sort_keys=True)
    for k, o in sorted(out.items(), key=lambda kv: -kv[1]["time_us"] * kv[1]["launches"]):
        print(f"{k:40s} n={o['launches']:3d} t={o['time_us']:8.1f}us traffic={o['traffic_bytes'] / 1e6:8.1f}MB "
              f"tensor={o['tensor_active_pct']:5.1f}% lts={o['lts_pct']:5.1f}% dram={o['dram_pct']:5.1f}%")


if __name__ == "__main__":
    main()

"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum [+ dram__bytes_read/write.sum]) per kernel family.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv "<header line>" > profiles/launches_rNN_summary.txt"""
import collections
import csv
import re
import sys


def family(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"^bgp::", "", name)
    m = re.match(r"gemm_nt_kernel<(\d+), (\d+)", name)
    if m:
        return f"gemm_nt_kernel<{m.group(1)}x{m.group(2)}>"
    return re.split(r"[<(]", name)[0]


def main():
    traffic_json = None
    if "--traffic-json" in sys.argv:
        i = sys.argv.index("--traffic-json")
        traffic_json = sys.argv[i + 1]
        del sys.argv[i:i + 2]
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = f.readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    for r in csv.DictReader(lines[start:]):
        rows.append(r)
    per = collections.defaultdict(lambda: {"ids": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
    for r in rows:
        fam = per[family(r["Kernel Name"])]
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Name"] == "gpu__time_duration.sum":
            fam["ids"].add(r["ID"])
            fam["ns"] += v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r["Metric Unit"], 1.0)
        elif r["Metric Name"] == "dram__bytes_read.sum":
            fam["rd"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
        elif r["Metric Name"] == "dram__bytes_write.sum":
            fam["wr"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
    total = sum(f["ns"] for f in per.values())
    for h in sys.argv[2:]:
        print("# " + h)
    print("# per-launch times are cold-cache and serialised under ncu: compare SHARES only")
    print(f"total_kernel_time_us {total / 1e3:.0f}   dram_read_GB {sum(f['rd'] for f in per.values()) / 1e9:.1f}   "
          f"dram_write_GB {sum(f['wr'] for f in per.values()) / 1e9:.1f}   launches {sum(len(f['ids']) for f in per.values())}")
    for name, f in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
        n = max(1, len(f["ids"]))
        print(f"{name:<34} launches={n:>5} total_ms={f['ns'] / 1e6:>9.2f} share={100 * f['ns'] / total:>5.1f}% "
              f"avg_us={f['ns'] / n / 1e3:>9.1f} dram_read_GB={f['rd'] / 1e9:>8.2f} dram_write_GB={f['wr'] / 1e9:>8.2f}")


    if traffic_json:
        import json
        out = {name: {"dram_bytes_per_launch": (f["rd"] + f["wr"]) / max(1, len(f["ids"])), "launches": len(f["ids"]),
                      "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum averaged over the kernel's launches in " + path}
               for name, f in per.items()}
        with open(traffic_json, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()

"""ncu launch-list CSV (long format, one row per kernel x metric) -> markdown table + per-kernel totals.
Usage: python tools/ncu_launches_md.py launches.csv "title / command line" > profiles/launches_rN.md
Prints the summed DRAM traffic of the tcgen05 conv launches on stderr (for profiles/traffic.json)."""
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("rnb::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = name.replace("unnamed>::", "").replace("(bool)", "")
    depth = 0
    for i, ch in enumerate(name):  # cut the argument list: first "(" outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            name = name[:i]
            break
    return name[:70]


def main(path, title):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    per = {}
    order = []
    for r in rd:
        k = int(r["ID"])
        if k not in per:
            per[k] = {"name": short(r["Kernel Name"])}
            order.append(k)
        per[k][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        per[k]["unit:" + r["Metric Name"]] = r["Metric Unit"]

    def us(d):
        v = d.get("gpu__time_duration.sum", 0.0)
        u = d.get("unit:gpu__time_duration.sum", "ns")
        return v / 1e3 if u in ("ns", "nsecond") else (v if u.startswith("u") else v * 1e3)

    def mb(d, key):
        v = d.get(key, 0.0)
        u = d.get("unit:" + key, "byte")
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)

    tot = sum(us(per[k]) for k in order)
    print(f"# ncu launch list — `{title}`\n")
    print("Times are cold-cache and serialised (profiler replay): compare SHARES, not absolutes.\n")
    print("| # | kernel | time us | share % | DRAM rd MB | DRAM wr MB | DRAM % | tensor pipe % |")
    print("|---|---|---|---|---|---|---|---|")
    agg = {}
    conv_traffic = 0.0
    for i, k in enumerate(order):
        d = per[k]
        t = us(d)
        rdmb, wrmb = mb(d, "dram__bytes_read.sum"), mb(d, "dram__bytes_write.sum")
        dram = d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
        tens = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        print(f"| {i} | `{d['name']}` | {t:.1f} | {100 * t / tot:.1f} | {rdmb:.0f} | {wrmb:.0f} | {dram:.0f} | {tens:.0f} |")
        base = d["name"].split("<")[0]
        a = agg.setdefault(base, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += rdmb + wrmb
        if any(s in base for s in ("conv_igemm", "conv3x3_halo", "bneck_")):
            conv_traffic += (rdmb + wrmb) * 1e6
    print("\n## Per-kernel totals\n")
    print("| kernel | launches | time us | share % | DRAM traffic MB |")
    print("|---|---|---|---|---|")
    for base, a in agg.items():
        print(f"| `{base}` | {a[0]} | {a[1]:.0f} | {100 * a[1] / tot:.1f} | {a[2]:.0f} |")
    print(f"| total | {len(order)} | {tot:.0f} | 100 | {sum(a[2] for a in agg.values()):.0f} |")
    print(f"conv traffic bytes: {conv_traffic:.0f}", file=sys.stderr)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")

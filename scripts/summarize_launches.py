"""Per-kernel summary of an ncu launch list (ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--clock-control none --csv --log-file <file>): launches, total / mean device time, share, DRAM bytes read / written."""
import collections
import csv
import io
import re
import sys


def load(path):
    import gzip
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        lines = [l for l in f.read().splitlines() if not l.startswith("==")]
    return list(csv.DictReader(io.StringIO("\n".join(lines))))


def main(path, out=None):
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in load(path):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        a = agg[name]
        if r["Metric Name"] == "gpu__time_duration.sum":
            a[0] += 1
            a[1] += {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(r["Metric Unit"], 1e-3) * v
        elif r["Metric Name"] == "dram__bytes_read.sum":
            a[2] += v * mult.get(r["Metric Unit"], 1)
        elif r["Metric Name"] == "dram__bytes_write.sum":
            a[3] += v * mult.get(r["Metric Unit"], 1)
    tot = sum(a[1] for a in agg.values())
    f = open(out, "w") if out else sys.stdout
    print("kernel,launches,total_us,share_pct,mean_us,dram_read_GB,dram_write_GB", file=f)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k},{a[0]},{a[1]:.1f},{100 * a[1] / tot:.2f},{a[1] / max(a[0], 1):.2f},{a[2] / 1e9:.3f},{a[3] / 1e9:.3f}", file=f)
    print(f"TOTAL,{sum(a[0] for a in agg.values())},{tot:.1f},100.00,,{sum(a[2] for a in agg.values()) / 1e9:.3f},{sum(a[3] for a in agg.values()) / 1e9:.3f}", file=f)


if __name__ == "__main__":
    main(*sys.argv[1:3])

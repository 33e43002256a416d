"""Summarise ncu output for profiles/: per-kernel table from a `--set full` report (.ncu-rep) or a
launch list CSV (--metrics gpu__time_duration.sum).

    python tools/ncu_summary.py report.ncu-rep > profiles/rNN_xxx.txt
    python tools/ncu_summary.py launches.csv  > profiles/rNN_launches.txt
    python tools/ncu_summary.py xxx_full_raw.csv > profiles/rNN_xxx.txt      (raw page exported on the GPU box)
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
]


def short(name):
    return name.split("(")[0].replace("ecgmm::", "").replace("void ", "")[:44]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def from_report(path):
    if path.endswith(".csv"):  # `ncu -i x.ncu-rep --page raw --csv` already run on the GPU box (tools/ncu_step.sh)
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {k: hdr.index(k) for k, _ in KEYS if k in hdr}
    iname = hdr.index("Kernel Name")
    print(f"# {path}: one row per profiled launch (ncu --set full --clock-control none; cold-cache, serialised)")
    print(f"{'kernel':44s} {'time_us':>9s} {'dram_rd_MB':>10s} {'dram_wr_MB':>10s} {'dram_%':>7s} {'l2_%':>6s} "
          f"{'tensor_%':>8s} {'sm_%':>6s} {'regs':>5s} {'grid':>6s}")
    agg = collections.OrderedDict()
    for r in rows[2:]:
        if len(r) <= iname:
            continue
        g = lambda k: r[idx[k]] if k in idx else "0"  # noqa: E731
        t = to_us(g("gpu__time_duration.sum"), units[idx["gpu__time_duration.sum"]])
        rd = to_bytes(g("dram__bytes_read.sum"), units[idx["dram__bytes_read.sum"]]) / 1e6
        wr = to_bytes(g("dram__bytes_write.sum"), units[idx["dram__bytes_write.sum"]]) / 1e6
        vals = [float(g(k).replace(",", "") or 0) for k, _ in KEYS[3:]]
        print(f"{short(r[iname]):44s} {t:9.1f} {rd:10.2f} {wr:10.2f} {vals[0]:7.1f} {vals[1]:6.1f} {vals[2]:8.1f} "
              f"{vals[3]:6.1f} {int(vals[4]):5d} {int(vals[5]):6d}")
        a = agg.setdefault(short(r[iname]), [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += rd + wr
        a[3] += vals[2] * t
        a[4] += vals[0] * t
    print("\n# per kernel: launches, total us, total DRAM MB, time-weighted tensor-pipe %, time-weighted DRAM %")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} n={a[0]:3d} {a[1]:10.1f} us {a[2]:10.1f} MB  tensor {a[3] / a[1]:5.1f}%  dram {a[4] / a[1]:5.1f}%")


def from_launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        t = to_us(row["Metric Value"], row["Metric Unit"])
        a = agg[short(row["Kernel Name"])]
        a[0] += 1
        a[1] += t
        tot += t
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us in total "
          "(ncu gpu__time_duration.sum, --clock-control none; cold-cache and serialised: compare SHARES)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} n={n:4d} {t:10.1f} us {100 * t / tot:5.1f}%")


if __name__ == "__main__":
    p = sys.argv[1]
    (from_report if (p.endswith(".ncu-rep") or p.endswith("_raw.csv")) else from_launches)(p)

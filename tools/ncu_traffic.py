"""profiles/r02_traffic.json from the per-class `ncu --set full` captures of tools/ncu_traffic_r02.sh.

    python tools/ncu_traffic.py gpurun_out/r02z > profiles/r02_ncu_class_traffic.txt     (also writes the JSON)

Each capture holds ONE training step's launches of one roofline class of bench.py (third step at per-GPU batch 64,
selected by NVTX range).  Per class: launches, device time, DRAM bytes read + written (dram__bytes_read.sum +
dram__bytes_write.sum), time-weighted tensor-pipe and DRAM utilisation.  bench.py scales dram_bytes_per_step by
batch / 64 and divides by the class's calls per step to report roofline.traffic per launch."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = ["conv_fwd", "conv_dgrad", "conv_wgrad", "bn_bwd_apply", "bn_bwd_reduce", "bn_apply", "perturb_fused"]
UNIT_B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
UNIT_T = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def load(path):
    rows = list(csv.reader(open(path).read().splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                      "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                                      "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") if k in hdr}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        f = lambda k: float(r[col[k]].replace(",", "") or 0)  # noqa: E731
        out.append({"name": r[col["Kernel Name"]].split("(")[0].replace("ecgmm::", "").replace("void ", ""),
                    "us": f("gpu__time_duration.sum") * UNIT_T.get(units[col["gpu__time_duration.sum"]], 1),
                    "rd": f("dram__bytes_read.sum") * UNIT_B.get(units[col["dram__bytes_read.sum"]], 1),
                    "wr": f("dram__bytes_write.sum") * UNIT_B.get(units[col["dram__bytes_write.sum"]], 1),
                    "tensor": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    "dram": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")})
    return out


def main():
    prefix = sys.argv[1]
    result = {"source": f"{os.path.basename(prefix)}_full_<class>.csv: ncu --set full --clock-control none, the launches of "
                        "one training step (third step of tools/one_step.py at per-GPU batch 64) per roofline class, selected "
                        "by the NVTX ranges of ecgmm.ops (tools/ncu_traffic_r02.sh); dram__bytes_read.sum + dram__bytes_write.sum",
              "per_gpu_batch": 64, "classes": {}}
    for c in CLASSES:
        path = f"{prefix}_full_{c}.csv"
        if not os.path.exists(path):
            continue
        ls = load(path)
        if not ls:
            continue
        t = sum(x["us"] for x in ls)
        by = sum(x["rd"] + x["wr"] for x in ls)
        print(f"# {c}: {len(ls)} launches, {t:.1f} us, DRAM {by / 1e6:.1f} MB "
              f"(read {sum(x['rd'] for x in ls) / 1e6:.1f} / write {sum(x['wr'] for x in ls) / 1e6:.1f}), "
              f"tensor pipe {sum(x['tensor'] * x['us'] for x in ls) / t:.1f} %, DRAM {sum(x['dram'] * x['us'] for x in ls) / t:.1f} % (time-weighted)")
        for x in ls:
            print(f"   {x['name'][:52]:52s} {x['us']:9.1f} us  rd {x['rd'] / 1e6:9.2f} MB  wr {x['wr'] / 1e6:9.2f} MB  "
                  f"tensor {x['tensor']:5.1f} %  dram {x['dram']:5.1f} %")
        result["classes"][c] = {"dram_bytes_per_step": by, "kernels_profiled": len(ls), "device_us": round(t, 1)}
    json.dump(result, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

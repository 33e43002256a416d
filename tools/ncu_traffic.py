"""profiles/r02_traffic.json + a per-class table from the light ncu pass of tools/ncu_traffic_r02.sh.

    python tools/ncu_traffic.py gpurun_out/r02z_step_metrics.csv > profiles/r02z_ncu_class_traffic.txt   (also writes the JSON)

Input: `ncu --profile-from-start off --metrics <time, dram read, dram write, tensor pipe %, dram %> --csv` over every
launch of ONE training step (the third step of tools/one_step.py at per-GPU batch 64, ECGMM_SIDE_STREAM=0).  Launches
are assigned to bench.py's roofline classes by kernel name; the conv forward and data-gradient classes share kernels and
are told apart by order (every forward launch of a step precedes every backward launch: the first 28 launches of the
igemm_nt family are the forward).  Per class: launches, device time, DRAM bytes read + written
(dram__bytes_read.sum + dram__bytes_write.sum), time-weighted tensor-pipe and DRAM utilisation.  bench.py scales
dram_bytes_per_step by batch / 64 and divides by the class's calls per step to report roofline.traffic per launch."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT_B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
UNIT_T = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}
FWD_LAUNCHES = 28  # conv forward calls per training step of the fusion model (kernels.conv_fwd.launches of the bench line)


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    out = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = out.setdefault(row["ID"], {"name": row["Kernel Name"].split("(")[0].replace("ecgmm::", "").replace("void ", "")})
        v = float(row["Metric Value"].replace(",", "") or 0)
        m, u = row["Metric Name"], row["Metric Unit"]
        if m == "gpu__time_duration.sum":
            d["us"] = v * UNIT_T.get(u, 1)
        elif m == "dram__bytes_read.sum":
            d["rd"] = v * UNIT_B.get(u, 1)
        elif m == "dram__bytes_write.sum":
            d["wr"] = v * UNIT_B.get(u, 1)
        elif m.startswith("sm__pipe_tensor"):
            d["tensor"] = v
        elif m.startswith("gpu__dram_throughput"):
            d["dram"] = v
    return list(out.values())


def classify(ls):
    nt_seen = 0
    for x in ls:
        n = x["name"]
        if n.startswith("igemm_nt"):
            x["cls"] = "conv_fwd" if nt_seen < FWD_LAUNCHES else "conv_dgrad"
            nt_seen += 1
        elif n.startswith("wgrad_halo") or n.startswith("igemm_tn") or n.startswith("tn_reduce"):
            x["cls"] = "conv_wgrad"
        elif n.startswith("bn_bwd_apply") or n.startswith("stem_bwd_apply"):
            x["cls"] = "bn_bwd_apply"
        elif n.startswith("bn_bwd_reduce"):
            x["cls"] = "bn_bwd_reduce"
        elif n.startswith("bn_apply"):
            x["cls"] = "bn_apply"
        elif n.startswith("bn_relu_maxpool"):
            x["cls"] = "bn_pool_fwd"
        elif n.startswith("stem_fwd_ring"):
            x["cls"] = "stem_fwd"
        elif n.startswith("stem_wgrad"):
            x["cls"] = "stem_wgrad"
        elif n.startswith("chan_stats"):
            x["cls"] = "bn_stats"
        else:
            x["cls"] = "other"
    return ls


def main():
    path = sys.argv[1]
    ls = classify([x for x in launches(path) if "us" in x])
    total = sum(x["us"] for x in ls)
    result = {"source": f"{os.path.basename(path)}: ncu --profile-from-start off --metrics (time, dram__bytes_read.sum, "
                        "dram__bytes_write.sum, tensor pipe %, dram %) --clock-control none over every launch of the third "
                        "training step of tools/one_step.py at per-GPU batch 64 (tools/ncu_traffic_r02.sh; the shipping build: "
                        "tools/r02_call_gg.sh), classified by "
                        "tools/ncu_traffic.py", "per_gpu_batch": 64, "classes": {}}
    print(f"# {path}: {len(ls)} launches of one training step at batch 64, {total:.0f} us of device time "
          "(cold-cache, serialised under ncu: compare shares)")
    agg = collections.OrderedDict()
    for x in ls:
        a = agg.setdefault(x["cls"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "t": 0.0, "d": 0.0, "k": collections.OrderedDict()})
        a["n"] += 1
        a["us"] += x["us"]
        a["rd"] += x.get("rd", 0)
        a["wr"] += x.get("wr", 0)
        a["t"] += x.get("tensor", 0) * x["us"]
        a["d"] += x.get("dram", 0) * x["us"]
        k = a["k"].setdefault(x["name"], [0, 0.0, 0.0])
        k[0] += 1
        k[1] += x["us"]
        k[2] += x.get("rd", 0) + x.get("wr", 0)
    print(f"{'class':16s} {'launches':>8s} {'us':>10s} {'share':>7s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s} {'tensor %':>9s} {'DRAM %':>7s}")
    for c, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{c:16s} {a['n']:8d} {a['us']:10.1f} {a['us'] / total:7.3f} {a['rd'] / 1e6:11.1f} {a['wr'] / 1e6:11.1f} "
              f"{a['t'] / a['us']:9.1f} {a['d'] / a['us']:7.1f}")
        if c != "other":
            result["classes"][c] = {"dram_bytes_per_step": a["rd"] + a["wr"], "kernels_profiled": a["n"],
                                    "device_us": round(a["us"], 1)}
    print("\n# per kernel within a class: launches, us, DRAM MB")
    for c, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        for k, v in sorted(a["k"].items(), key=lambda kv: -kv[1][1])[:12]:
            print(f"{c:16s} {k[:56]:56s} n={v[0]:3d} {v[1]:9.1f} us {v[2] / 1e6:9.1f} MB")
    json.dump(result, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

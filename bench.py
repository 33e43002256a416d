"""Benchmark of the fusion-classifier training step (BASELINE.json: "train samples/sec,
ResNet18+signal+clinical fusion at 1/2/4/8 B200; % roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--global-batch B]

Workload (configs[2] of BASELINE.json): the full multimodal training step -- forward, CE +
0.1*var_loss, backward, Adam, and for N > 1 the bucketed gradient all-reduce -- at native
3x250x2500 images (bf16), 2476-sample single-lead signals, 24 clinical features, GLOBAL batch
512 sharded over the N GPUs of one node (strong scaling: 512/N samples per GPU), synthetic
data, random-init weights.  One process per GPU (torchrun for N > 1).

One JSON line on rank 0:
  value     whole-job samples/s with the step's inputs already resident in HBM
  e2e       same, through the public API with HOST (pinned) inputs: the H2D copy of every
            step's batch and the D2H read of its loss are inside the timed region
  roofline  the dominant kernel class, timed live with CUDA events on the launching stream
  cpu_baseline  the CPU oracle (== reference code path) on a bounded sample, rank 0, N=1 only
  --impl reference : the oracle timed on the host cores with the same metric/config
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, L, F = 250, 2500, 2476, 24
GLOBAL_BATCH = 512
CONV_GFLOP_TRAIN_PER_SAMPLE = 135.642
CONFIG_NAME = "configs[2]: full multimodal training step, 3x250x2500 bf16 images, global batch 512, data-parallel"


def synth_batch(B, seed, pinned=False, image_dtype="uint8"):
    """SURVEY.md section 8d cfg3 inputs.  image_dtype: what the loader hands over -- "uint8" (default: the 8-bit pixels
    the reference's ToTensor + Normalize(0.5, 0.5) starts from, dataset.py:119-123; the normalisation runs inside the
    first kernel, bit-identical to the host transform), "fp32" (the reference loader's own output) or "bf16"."""
    import torch

    g = torch.Generator().manual_seed(seed)
    image = torch.randn(B, 3, H, W, generator=g).clamp_(-1, 1)
    if image_dtype == "uint8":
        image = image.mul_(127.5).add_(127.5).round_().clamp_(0, 255).to(torch.uint8)
    elif image_dtype == "bf16":
        image = image.to(torch.bfloat16)
    ecg = torch.randn(B, L, generator=g)
    clin = torch.randn(B, F, generator=g)
    labels = (torch.rand(B, generator=g) < 0.4).long()  # 88/220 abnormal in the reference data set
    out = [image, ecg, clin, labels]
    if pinned:
        out = [t.pin_memory() for t in out]
    return out


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference_steps(steps, warmup, batch):
    """Times the oracle (bit-identical restatement of the reference, oracle/model.py) on the host
    cores: full train step at 3x250x2500 on a bounded batch.  Returns (samples/s, ms/step, cores)."""
    import torch

    from oracle import model as om

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    m = om.ECGMultimodalModel()
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    image, ecg, clin, labels = synth_batch(batch, 42, image_dtype="fp32")
    for _ in range(warmup):
        om.fusion_train_step(m, opt, image, ecg, clin, labels)
    t0 = time.perf_counter()
    for _ in range(steps):
        om.fusion_train_step(m, opt, image, ecg, clin, labels)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return batch / dt, dt * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 4
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    sps, ms, cores = cpu_reference_steps(steps, warmup, batch)
    line = {
        "impl": "reference", "metric": "train samples/sec", "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CONFIG_NAME, "image": [3, H, W], "signal_len": L, "clinical_features": F,
                   "global_batch": GLOBAL_BATCH, "step_sample": f"batch {batch} per step on the host CPU"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} full train steps (fwd+loss+bwd+Adam) of batch {batch} at 3x{H}x{W}, fp32, "
                                   f"oracle/model.py (bit-identical to the reference module), torch {cores} threads"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    if not os.path.exists(os.path.join(ROOT, "ecg-multimodal-model_b200", "libecgmm.so")):
        # (git-ignored artefact) build once; under torchrun only local rank 0 builds, the others wait for the file
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            import __graft_entry__ as g

            g.build()
        else:
            for _ in range(600):
                if os.path.exists(os.path.join(ROOT, "ecg-multimodal-model_b200", "libecgmm.so")):
                    break
                time.sleep(1.0)
            time.sleep(2.0)
    import ecgmm
    from ecgmm import lib, ops
    from ecgmm import nn as enn
    from ecgmm import optim as eoptim
    from ecgmm.parallel import DataParallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    lib.require_device()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gb = args.global_batch
    if gb % world:
        raise SystemExit(f"global batch {gb} not divisible by {world} GPUs")
    B = gb // world
    dev = torch.device("cuda", local)

    class Cfg:
        num_classes = 2
        device = dev

    torch.manual_seed(42)
    model = ecgmm.ECGMultimodalModel(Cfg)
    model.train()
    dp = DataParallel(model) if world > 1 else None
    net = dp if dp is not None else model
    crit = enn.CrossEntropyLoss()
    opt = eoptim.Adam(model.parameters(), lr=1e-4)

    host = synth_batch(B, 42 + rank, pinned=True, image_dtype=args.image_dtype)
    resident = [t.to(dev) for t in host]

    def step(batch):
        image, ecg, clin, labels = batch
        opt.zero_grad()
        out = net(image, ecg, clin)
        loss = crit(out[3], labels) + 0.1 * out[4]
        loss.backward()
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {}

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time by CUDA events; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for _ in range(steps):
            fn()
        host_ms["last"] = (time.perf_counter() - t_host) * 1e3 / steps  # CPU time to ENQUEUE one step
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up (also builds weight shadows, Adam state and the allocator's working set)
    for _ in range(args.warmup):
        step(resident)
    torch.cuda.synchronize()

    # ---- the step as ONE CUDA graph (ecgmm.graph.GraphedTrainStep: same kernels, one launch per step); the eager
    # loop stays available with --launch eager and is what the per-kernel profile below runs
    launch, graph_note = "eager", None
    if args.launch == "graph":
        gstep = None
        try:
            from ecgmm.graph import GraphedTrainStep

            gstep = GraphedTrainStep(net, crit, opt, resident, restore=False)
        except Exception as ex:  # capture is an optimisation of the HOST side only; say so and carry on eagerly
            graph_note = f"{type(ex).__name__}: {ex}"[:300]
            torch.cuda.synchronize()
        if world > 1:  # every rank replays the graph or none does (the collectives inside have to pair up)
            okf = torch.tensor([1 if gstep is not None else 0], device=dev)
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
            if int(okf.item()) == 0 and gstep is not None:
                gstep, graph_note = None, "capture failed on another rank"
        if gstep is not None:
            for _ in range(args.warmup):
                gstep(*gstep.inputs)
            torch.cuda.synchronize()
            eager_step = step

            def step(batch):  # noqa: F811
                return gstep(*batch)

            resident = gstep.inputs
            launch = "cuda_graph"

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    # ---- device-resident throughput
    launches0 = lib.launch_count()
    total_ms = timed(lambda: step(resident), args.steps)
    launches = lib.launch_count() - launches0
    ms_per_step = total_ms / args.steps
    value = gb / (ms_per_step / 1e3)
    host_enqueue_ms = host_ms["last"]

    # ---- end to end through the public API: the loop of train.py:60-86 over ecgmm.data.Prefetcher (pinned host batch
    # -> device every step on a copy stream, double-buffered, overlapping the previous step) + the loss read back every
    # step; every copy and every read-back happens inside the timed region.
    from ecgmm.data import Prefetcher

    class HostLoader:  # a "DataLoader" that yields the same pinned host batch n times
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __iter__(self):
            return iter([host] * self.n)

    prefetch = Prefetcher(HostLoader(0), dev, depth=2)
    last = {}
    state = {"i": 0}

    # The loss of step i is copied D2H (pinned scalar, enqueued right behind the step) and READ by the host while step
    # i+1 runs: every step's loss is read inside the timed region, but the host never sits between two steps waiting
    # for a scalar while the GPU idles through the next launch (train.py:84 reads loss.item() at once; a loop that logs
    # one step late is the same training).  ECGMM_BENCH_SYNC_LOSS=1 restores the blocking read.
    sync_loss = os.environ.get("ECGMM_BENCH_SYNC_LOSS", "0") == "1"
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_step(batch):
        i = state["i"]
        slot = i & 1
        loss = step(batch)
        if sync_loss:
            last["loss"] = float(loss.item())  # D2H read of the step's result
        else:
            loss_host[slot].copy_(loss.detach().float(), non_blocking=True)
            loss_ev[slot].record()
            if i > 0:
                loss_ev[slot ^ 1].synchronize()
                last["loss"] = float(loss_host[slot ^ 1])
        state["i"] = i + 1

    def e2e_run(steps):
        state["i"] = 0
        prefetch.loader = HostLoader(steps)
        for batch in prefetch:
            e2e_step(batch)
        if not sync_loss and steps > 0:  # the last step's loss
            s_last = (state["i"] - 1) & 1
            loss_ev[s_last].synchronize()
            last["loss"] = float(loss_host[s_last])

    e2e_note = None
    try:
        e2e_run(2)
    except Exception as ex:  # never lose the bench line to the pipelined read: fall back to the blocking one
        e2e_note = f"pipelined loss read failed ({type(ex).__name__}: {ex}); blocking read used"[:200]
        sync_loss = True
        torch.cuda.synchronize()
        e2e_run(2)
    h2d0 = prefetch.h2d_bytes
    e2e_ms = timed(lambda: e2e_run(args.steps), 1) / args.steps
    h2d_per_step = (prefetch.h2d_bytes - h2d0) / args.steps
    e2e_value = gb / (e2e_ms / 1e3)
    clocks = sampler.stop() if sampler else None

    # ---- per-kernel-class device times over one more (eager) step (CUDA events on the launching stream)
    if launch == "cuda_graph":
        step = eager_step
    # (the signal / clinical branches normally run on a side stream underneath the image encoder; for this pass they
    # are serialised onto the main stream, otherwise an event pair around a small side-stream kernel also measures the
    # time it waited for SMs and pollutes its class)
    overlap, model.overlap_branches = model.overlap_branches, False
    ops.PROFILE = []
    step(resident)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    model.overlap_branches = overlap
    n0 = lib.launch_count()
    step(resident)  # kernels of one step, counted by the library (a graph replay re-issues exactly these)
    torch.cuda.synchronize()
    launches_per_step_eager = lib.launch_count() - n0
    classes, detail = {}, {}
    for kind, work, a, b, *_ in prof:
        ms = a.elapsed_time(b)
        for table, key in ((classes, kind.split("/")[0]), (detail, kind)):
            c = table.setdefault(key, {"launches": 0, "ms": 0.0, "work": 0.0})
            c["launches"] += 1
            c["ms"] += ms
            c["work"] += work
    if rank == 0 and args.detail:
        for k, c in sorted(detail.items(), key=lambda kv: -kv[1]["ms"]):
            tensor = k.startswith("conv") or k.startswith("stem")
            ach = c["work"] / (c["ms"] * 1e-3) / (1e12 if tensor else 1e9)
            print(f"# {k:34s} n={c['launches']:3d} {c['ms']:8.3f} ms {ach:9.1f} {'TFLOP/s' if tensor else 'GB/s'}",
                  file=sys.stderr)

    def shutdown():
        """Leave the process group without hanging.  Observed on 8 GPUs: after every rank had finished and rank 0 had
        printed its line, dist.destroy_process_group() did not return while a captured CUDA graph containing NCCL
        kernels was alive.  So: drain the device, meet the other ranks, drop the graph, then give the teardown 15 s
        in a helper thread and leave the process hard either way (all results are out by then)."""
        if world <= 1:
            return
        sys.stdout.flush()
        sys.stderr.flush()
        # whatever blocks below (a peer that already died, a communicator that will not drain): every result is out,
        # so the process leaves after 60 s at the latest instead of holding the launcher until its own limit
        wd = threading.Timer(60.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        try:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            if launch == "cuda_graph":
                gstep.graph.reset()
            t = threading.Thread(target=dist.destroy_process_group, daemon=True)
            t.start()
            t.join(15.0)
        except Exception as ex:  # the measurement is complete and printed: a teardown problem must not fail the run
            print(f"# teardown: {type(ex).__name__}: {ex}", file=sys.stderr)
        finally:
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)  # also skips interpreter finalisation, where the NCCL / graph destructors could block again

    if rank != 0:
        shutdown()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peaks = json.load(open(peaks_path))
        tf_peak, hbm_peak, src = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained)"
    else:
        tf_peak, hbm_peak, src = 1400.0, 6650.0, "fallback (B200_PROFILING.md)"
    kernels = {}
    for kind, c in classes.items():
        tensor = kind.startswith("conv") or kind.startswith("stem")
        ach = c["work"] / (c["ms"] * 1e-3) / (1e12 if tensor else 1e9) if c["ms"] > 0 else 0.0
        kernels[kind] = {"bound": "tensor" if tensor else "hbm", "launches": c["launches"],
                         "ms_per_step": round(c["ms"], 4), "achieved": round(ach, 2),
                         "unit": "TFLOP/s" if tensor else "GB/s",
                         "frac": round(ach / (tf_peak if tensor else hbm_peak), 4)}
    # The 7x7 stem is HBM-bound (arithmetic intensity 124 flop/B): next to its tensor figure,
    # report the bytes it has to move -- the space-to-depth image (32 B per 2x2 pixel block) and the 64-channel output
    # (128 B per output pixel), once each -- against the HBM peak.
    try:
        stem_bytes = float(B) * (((H + 1) // 2) * ((W + 1) // 2) * 32 + ((H - 1) // 2 + 1) * ((W - 1) // 2 + 1) * 128)
        for kind in ("stem_fwd", "stem_wgrad"):
            if kind in kernels and classes[kind]["ms"] > 0:
                gbs = stem_bytes / (classes[kind]["ms"] * 1e-3) / 1e9
                kernels[kind]["hbm"] = {"achieved": round(gbs, 1), "unit": "GB/s", "frac": round(gbs / hbm_peak, 4)}
    except Exception:
        pass
    dom = max(classes, key=lambda k: classes[k]["ms"]) if classes else None
    roofline = None
    if dom:
        k = kernels[dom]
        kname = {"conv_wgrad": "wgrad_halo_kernel<128,T> / <64> + igemm_tn_kernel + their split-K folds (conv weight-gradient)",
                 "conv_fwd": "igemm_nt_pair_kernel + igemm_nt_stack_kernel + igemm_nt_kernel (conv forward)",
                 "conv_dgrad": "igemm_nt_pair_kernel + igemm_nt_stack_kernel + igemm_nt_kernel (conv data-gradient)",
                 "bn_bwd_apply": "bn_bwd_apply_fast_kernel / bn_bwd_apply_kernel + stem_bwd_apply_rows_kernel (BatchNorm backward, dx)"}.get(dom, dom)
        # DRAM bytes of this class from the committed `ncu --set full` capture of one training step at per-GPU batch 64
        # (profiles/r02_traffic.json, tools/ncu_traffic_r02.sh), scaled to this run's per-GPU batch, per call of the class
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            ent = tj.get("classes", {}).get(dom)
            if ent and k["launches"]:
                traffic = ent["dram_bytes_per_step"] * B / tj["per_gpu_batch"] / k["launches"]
                traffic_src = tj.get("source")
        roofline = {"kernel": kname, "bound": k["bound"], "achieved": k["achieved"],
                    "peak": tf_peak if k["bound"] == "tensor" else hbm_peak, "unit": k["unit"], "frac": k["frac"],
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": src,
                    "share_of_step": round(classes[dom]["ms"] / ms_per_step, 4), "launches_per_step": k["launches"]}
    conv_ms = sum(c["ms"] for kd, c in classes.items() if kd.startswith("conv") or kd.startswith("stem"))
    conv_flops = sum(c["work"] for kd, c in classes.items() if kd.startswith("conv") or kd.startswith("stem"))

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sps, ms, cores = cpu_reference_steps(3, 1, 4)
        cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"3 full train steps of batch 4 at 3x{H}x{W} fp32 through oracle/model.py ({ms:.0f} ms/step)"}

    line = {
        "metric": "train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": CONFIG_NAME, "image": [3, H, W], "signal_len": L, "clinical_features": F,
                   "global_batch": gb, "per_gpu_batch": B, "parallelism": f"dp{world}", "launch": launch,
                   "image_input": {"uint8": "uint8 pixels; ToTensor + Normalize(0.5, 0.5) fused into the first kernel",
                                   "fp32": "fp32 normalised tensors (the reference loader's output)",
                                   "bf16": "bf16 normalised tensors"}[args.image_dtype],
                   "l2": "per-step working set (>= 7 GB of activations per GPU) exceeds the 126 MB L2; no flush needed"},
        "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d_per_step * world, "api": "ecgmm.data.Prefetcher + the train.py loop body",
                "d2h_bytes_per_step": 4 * world, "last_loss": last.get("loss"),
                "loss_read": "blocking" if sync_loss else "pipelined: step i's loss read while step i+1 runs",
                "note": e2e_note},
        "gpu_launches": launches if launch == "eager" else launches_per_step_eager * args.steps,
        "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
        "gpu_launches_per_step": launches / args.steps if launch == "eager" else launches_per_step_eager,
        "graph_note": graph_note,
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels,
        "conv_total": {"ms_per_step": round(conv_ms, 3), "tflops": round(conv_flops / (conv_ms * 1e-3) / 1e12, 1) if conv_ms else None,
                       "frac_of_sustained_peak": round(conv_flops / (conv_ms * 1e-3) / 1e12 / tf_peak, 4) if conv_ms else None,
                       "share_of_step": round(conv_ms / ms_per_step, 4)},
        "cpu_baseline": cpu,
    }
    if dp is not None:
        line["allreduce"] = {"buckets_per_step": dp.buckets_last_step, "bytes_per_step": dp.bytes_last_step}
    print(json.dumps(line), flush=True)
    shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--image-dtype", default="uint8", choices=["uint8", "fp32", "bf16"],
                    help="what the loader hands over (host side of e2e and the resident batch alike)")
    ap.add_argument("--detail", action="store_true", help="per-shape kernel timings on stderr")
    ap.add_argument("--config", default="train", choices=["train", "signal", "perturb", "kfold"],
                    help="train: configs[2] of BASELINE.json, the headline (default); signal: configs[1] (12-lead 1D-CNN, "
                         "batch 256, 1 GPU); perturb: configs[3] (4096 masked variants per sample, samples sharded over "
                         "the GPUs); kfold: configs[4] (5 folds of a 10k-patient synthetic data set, folds sharded over the "
                         "GPUs).  The other configs print their own JSON line (tools/*_bench.py); extra arguments after "
                         "`--` are passed on to them")
    ap.add_argument("rest", nargs="*", help=argparse.SUPPRESS)
    ap.add_argument("--launch", default="graph", choices=["graph", "eager"],
                    help="graph: the training step replayed as one CUDA graph (ecgmm.graph); eager: kernel by kernel")
    args = ap.parse_args()
    if args.config != "train" and args.impl == "ours":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        mod = {"signal": "signal_bench", "perturb": "perturb_bench", "kfold": "kfold_bench"}[args.config]
        __import__(mod).main(list(args.rest))
        return
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""A/B of the stem / BatchNorm elementwise kernel variants (round 2, last GPU calls): per-kernel CUDA-event times at the
shapes of the training step, and bit-equality between the variants of one operation.

    python tools/elementwise_ab.py [--batch 64] [--iters 20]

Variants are selected per call through the library's environment switches:
    ECGMM_POOL_LEGACY / ECGMM_POOL_TILED_V1     bn_relu_maxpool: per-output kernel / tiled v1 / tiled branch-free (default)
    ECGMM_STEM_BWD_APPLY=0|1|2                  stem_bwd_apply: grid-stride / CTA per pooled row (regs) / (smem, default)
    ECGMM_BN_FAST=0                             bn_apply / bn_bwd_apply: generic kernels / register-resident coefficients (default)
    ECGMM_BN_ASYNC=0                            bn_bwd_apply / bn_bwd_reduce: loads into registers / through a cp.async ring (default)
(profiles/r02gg_ab128.txt also holds two variants that were measured and removed: coefficients in shared memory with a
register cap for 5 CTAs per SM, and four vectors per trip.)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ecgmm  # noqa: E402,F401
from ecgmm import lib, ops  # noqa: E402


def timed(fn, iters, flush):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)  # 512 MB > the 126 MB L2: every timed launch starts cold
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


class Env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only-bn", action="store_true", help="skip the stem sections")
    a = ap.parse_args()
    lib.require_device()
    dev = "cuda"
    torch.manual_seed(0)
    N, H, W, C = a.batch, 125, 1250, 64
    flush = torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device=dev)
    rep = {"batch": N, "shape": [H, W, C]}

    nb = torch.zeros((), dtype=torch.int64, device=dev)
    if not a.only_bn:
        x = torch.randn(N, H, W, C, device=dev).to(torch.bfloat16)
        gamma = (torch.randn(C, device=dev) * 0.5 + 0.2)  # both signs: the flipped-domain scan is exercised
        beta = torch.randn(C, device=dev) * 0.1
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        nb = torch.zeros((), dtype=torch.int64, device=dev)
        st = ops.bn_train_stats(x, gamma, beta, rm, rv, nb, 1e-5, 0.1)

        # ---- stem max-pool forward
        pool = {}
        outs = {}
        for name, env in (("per_output", dict(ECGMM_POOL_LEGACY="1")), ("tiled_v1", dict(ECGMM_POOL_TILED_V1="1")),
                          ("tiled_branchfree", {})):
            with Env(ECGMM_POOL_LEGACY=None, ECGMM_POOL_TILED_V1=None):
                with Env(**env):
                    outs[name] = ops.bn_relu_maxpool(x, st)
                    pool[name] = timed(lambda: ops.bn_relu_maxpool(x, st), a.iters, flush)
        ref = outs["per_output"]
        pool["bit_equal"] = all(torch.equal(o[0], ref[0]) and torch.equal(o[1], ref[1]) for o in outs.values())
        nbytes = 2.0 * x.numel() + 3.0 * ref[0].numel()
        pool["GBs"] = {k: nbytes / (v * 1e-3) / 1e9 for k, v in pool.items() if k != "bit_equal"}
        rep["bn_relu_maxpool_ms"] = pool
        y, arg = ref

        # ---- stem backward apply (+ the pooled-domain reduction and both finalize kernels in front of it)
        dy = torch.randn_like(y)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        bwd, douts = {}, {}
        for v in ("0", "1", "2"):
            with Env(ECGMM_STEM_BWD_APPLY=v):
                f = lambda: ops.bn_backward(x, dy, st, gamma, argmax=arg, pooled=y, beta=beta, dgamma=dg, dbeta=db)  # noqa: E731
                douts[v] = f()[0]
                ops.PROFILE = []
                for _ in range(a.iters):
                    flush.add_(1.0)
                    f()
                torch.cuda.synchronize()
                ts = sorted(e0.elapsed_time(e1) for k, _, e0, e1, _ in ops.PROFILE if k.startswith("bn_bwd_apply"))
                ops.PROFILE = None
                bwd[v] = ts[len(ts) // 2]
        bwd["bit_equal"] = all(torch.equal(douts[v], douts["0"]) for v in douts)
        nbytes = 4.0 * x.numel() + 3.0 * dy.numel()
        bwd["GBs"] = {k: nbytes / (v * 1e-3) / 1e9 for k, v in bwd.items() if k != "bit_equal"}
        rep["stem_bwd_apply_ms"] = bwd

    # ---- BatchNorm apply / backward apply / backward reduce: generic kernels (ECGMM_BN_FAST=0) against the fast paths
    bn = {}
    for (h, w, c) in ((63, 625, 64), (32, 313, 128), (16, 157, 256), (8, 79, 512)):
        xx = torch.randn(N, h, w, c, device=dev).to(torch.bfloat16)
        rr = torch.randn_like(xx)
        dd = torch.randn_like(xx)
        g2, b2 = torch.randn(c, device=dev), torch.randn(c, device=dev)
        s2 = ops.bn_train_stats(xx, g2, b2, None, None, None, 1e-5, 0.1)
        res, keep = {}, {}
        for tag, env in (("generic", dict(ECGMM_BN_FAST="0", ECGMM_BN_ASYNC="0")), ("fast", dict(ECGMM_BN_ASYNC="0")),
                         ("async", {})):
            with Env(ECGMM_BN_FAST=None, ECGMM_BN_ASYNC=None):
                with Env(**env):
                    y1, m1 = ops.bn_apply(xx, s2, relu=True, want_mask=True)
                    y2, m2 = ops.bn_apply(xx, s2, res=rr, relu=True, want_mask=True)
                    dgg, dbb = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
                    dx1, dz1 = ops.bn_backward(xx, dd, s2, g2, mask=m2, want_dz=True, dgamma=dgg, dbeta=dbb)
                    dx2, _ = ops.bn_backward(xx, dd, s2, g2, mask=m1, dgamma=dgg.clone(), dbeta=dbb.clone())
                    dx3, _ = ops.bn_backward(xx, dd, s2, g2, dgamma=dgg.clone(), dbeta=dbb.clone())
                    keep[tag] = (y1, m1, y2, m2, dx1, dz1, dx2, dx3, dgg.clone(), dbb.clone())
                    res[tag + "_apply_relu"] = timed(lambda: ops.bn_apply(xx, s2, relu=True, want_mask=True), a.iters, flush)
                    res[tag + "_apply_res_relu"] = timed(lambda: ops.bn_apply(xx, s2, res=rr, relu=True, want_mask=True),
                                                         a.iters, flush)
                    for kind, kw in (("bwd_mask_dz", dict(mask=m2, want_dz=True)), ("bwd_mask", dict(mask=m1))):
                        ops.PROFILE = []
                        for _ in range(a.iters):
                            flush.add_(1.0)
                            ops.bn_backward(xx, dd, s2, g2, dgamma=dgg, dbeta=dbb, **kw)
                        torch.cuda.synchronize()
                        for pref in ("bn_bwd_apply", "bn_bwd_reduce"):
                            ts = sorted(e0.elapsed_time(e1) for k, _, e0, e1, _ in ops.PROFILE if k.startswith(pref))
                            res[f"{tag}_{kind}_{pref[7:]}"] = ts[len(ts) // 2]
                        ops.PROFILE = None
        res["bit_equal"] = all(torch.equal(p, q) for t in ("fast", "async") for p, q in zip(keep["generic"], keep[t]))
        bn[f"C{c}"] = res
    rep["bn_ms"] = bn
    print(json.dumps(rep))


if __name__ == "__main__":
    main()

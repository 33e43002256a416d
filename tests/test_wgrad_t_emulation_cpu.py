"""CPU: index-level emulation of the transposed weight-gradient mode (csrc/conv_wgrad_halo.cu,
wgrad_halo_kernel<128, true> + make_plan + wgrad_halo_reduce_t_kernel): CTA types, split-K ranges, the operands each accumulator slot
multiplies (dY as A, one staged input row read at three pixel shifts as the N = 192 B operand), the workspace layout
[slot][192 columns][128 rows] and the reduction's decode back to dw[cout][cin][r][s] -- against torch's weight gradient.
The GPU parity of the kernel is in tests/test_conv_gpu.py; this pins the bookkeeping it was written from."""
import numpy as np
import pytest
import torch


def plan_t(N, H, W, Cin, Cout, sms):
    Ho, Wo = H, W
    best_kp, best_cost = 64, -1
    for kp in range(128, 31, -16):
        cost = -(-Wo // kp) * kp
        if best_cost < 0 or cost < best_cost:
            best_cost, best_kp = cost, kp
    q = dict(KP=best_kp, rps=1, tiles_w=-(-Wo // best_kp), Ho=Ho, Wo=Wo)
    q["total_kb"] = N * Ho * q["tiles_w"]
    q["cin_chunks"] = Cin // 64
    q["cin_pairs"] = (q["cin_chunks"] + 1) // 2
    q["cout_tiles"] = Cout // 128
    q["groupsA"] = q["cin_chunks"] * q["cout_tiles"]
    q["groupsB"] = q["cin_pairs"] * q["cout_tiles"]
    ks = max(1, sms // (q["groupsA"] + q["groupsB"]))
    q["ksA"] = q["ksB"] = min(ks, max(1, q["total_kb"]))
    q["wsB_off"] = q["groupsA"] * q["ksA"] * 2 * 192 * 128
    q["ws_floats"] = q["wsB_off"] + q["groupsB"] * q["ksB"] * 2 * 192 * 128
    return q


def emulate(x, dy, sms):
    """x [N,H,W,Cin], dy [N,H,W,Cout] float64 numpy -> dw [Cout][Cin][3][3] through the kernel's bookkeeping."""
    N, H, W, Cin = x.shape
    Cout = dy.shape[3]
    q = plan_t(N, H, W, Cin, Cout, sms)
    KP = q["KP"]
    ws = np.full(q["ws_floats"], np.nan)
    nA = q["groupsA"] * q["ksA"]
    grid = nA + q["groupsB"] * q["ksB"]
    xp = np.zeros((N, H + 2, W + 2 + KP, Cin))          # zero fill outside the image, as TMA does
    xp[:, 1:H + 1, 1:W + 1] = x
    dyp = np.zeros((N, H, W + KP, Cout))
    dyp[:, :, :W] = dy
    for bid_all in range(grid):
        typeB = bid_all >= nA
        bid = bid_all - nA if typeB else bid_all
        ksplit = q["ksB"] if typeB else q["ksA"]
        ks, g = bid % ksplit, bid // ksplit
        gdiv = q["cin_pairs"] if typeB else q["cin_chunks"]
        cc = 2 * (g % gdiv) if typeB else g % gdiv
        nt = g // gdiv
        pair_ok = typeB and cc + 1 < q["cin_chunks"]
        per = -(-q["total_kb"] // ksplit)
        kb0, kb1 = ks * per, min(q["total_kb"], ks * per + per)
        acc = np.zeros((2, 128, 192))                    # [slot][m = cout][n = s*64 + cin]
        for kb in range(kb0, kb1):
            twi, row = kb % q["tiles_w"], kb // q["tiles_w"]
            oh, img = row % q["Ho"], row // q["Ho"]
            w0 = twi * KP
            d = dyp[img, oh, w0:w0 + KP, nt * 128:(nt + 1) * 128]               # [KP pixels][128 cout]
            for i in range(2):
                r = 2 if typeB else i
                sl = (cc + i if pair_ok else cc) if (typeB and i == 1) else cc
                box = xp[img, oh + r, w0:w0 + KP + 2, sl * 64:(sl + 1) * 64]    # input row oh + r - 1, from pixel w0 - 1
                for s in range(3):
                    acc[i, :, s * 64:(s + 1) * 64] += d.T @ box[s:s + KP]
        off = (q["wsB_off"] if typeB else 0) + (g * ksplit + ks) * 2 * 192 * 128
        ws[off:off + 2 * 192 * 128] = acc.transpose(0, 2, 1).reshape(-1)        # [slot][column][row]
    assert not np.isnan(ws).any()                         # every slice written exactly by its CTA
    # ---- the reduction kernel's decode
    dw = np.zeros((Cout, Cin, 3, 3))
    for typeB in (False, True):
        groups = q["groupsB"] if typeB else q["groupsA"]
        ksplit = q["ksB"] if typeB else q["ksA"]
        base = q["wsB_off"] if typeB else 0
        part = ws[base:base + groups * ksplit * 2 * 192 * 128].reshape(groups, ksplit, 2, 192, 128).sum(1)
        for g in range(groups):
            for slot in range(2):
                if typeB:
                    r, chunk, nt = 2, 2 * (g % q["cin_pairs"]) + slot, g // q["cin_pairs"]
                    if chunk >= q["cin_chunks"]:
                        continue
                else:
                    r, chunk, nt = slot, g % q["cin_chunks"], g // q["cin_chunks"]
                for s in range(3):
                    # element (c = s*64 + cin_local, m = cout_local)
                    dw[nt * 128:(nt + 1) * 128, chunk * 64:(chunk + 1) * 64, r, s] += part[g, slot, s * 64:(s + 1) * 64, :].T
    return dw, q


@pytest.mark.parametrize("N,H,W,Cin,Cout,sms", [(1, 3, 40, 128, 128, 7), (2, 2, 150, 192, 128, 11), (1, 4, 33, 64, 256, 5),
                                                (1, 1, 20, 128, 256, 148)])
def test_transposed_wgrad_bookkeeping(N, H, W, Cin, Cout, sms):
    g = torch.Generator().manual_seed(N * 100 + W)
    x = torch.randn(N, H, W, Cin, generator=g, dtype=torch.float64)
    dy = torch.randn(N, H, W, Cout, generator=g, dtype=torch.float64)
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w, None, 1, 1)
    y.backward(dy.permute(0, 3, 1, 2))
    dw, q = emulate(x.numpy(), dy.numpy(), sms)
    assert q["ws_floats"] > 0 and np.abs(dw - w.grad.numpy()).max() <= 1e-9 * max(1.0, float(w.grad.abs().max()))

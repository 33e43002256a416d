"""CPU emulation of the CONTROL LOGIC of the rolling-accumulator kernel (csrc/conv_nt_stack.cu): the
weight-tile order, which output rows an input row feeds at the edges of a unit, the accumulator ring with its wrap
split, the hand-back discipline (a block is zero when first touched, complete when emitted).  The arithmetic of each
"MMA group" is a plain matmul here; what is checked is that the schedule of groups adds up to the convolution
(forward) and to its data gradient, for several unit lengths -- i.e. the part of the kernel that can be verified
without a GPU.  The formulas below are transcribed from the kernel line by line."""
import pytest
import torch
import torch.nn.functional as F

BLOCKS = 7  # kStBlocks


def tap_of_q(dgrad):
    # launch_nt_stack: q = shift*3 + k feeds output row i-1+k
    return [(k * 3 + (2 - sh)) if dgrad else ((2 - k) * 3 + sh) for sh in range(3) for k in range(3)]


def emulate(x, w_tiles, dgrad, seg):
    """x: [H, W, 64] input of ONE image (forward: activations, dgrad: dy); w_tiles: [9][64 n][64 k] in tap order
    t = r*3+s of the layout the kernel is given (w_fwd [O][R][S][I] / w_dgrad [I][R][S][O]).  Returns [H, W, 64]."""
    H, W, C = x.shape
    tq = tap_of_q(dgrad)
    sW = torch.stack([w_tiles[tq[q]] for q in range(9)])          # smem order
    out = torch.zeros(H, W, C)
    tmem = torch.zeros(BLOCKS, W, C)                               # one "strip" as wide as the image: rows = pixels
    touched = [False] * BLOCKS
    g0 = 0
    xpad = F.pad(x.permute(2, 0, 1), (1, 1, 0, 0)).permute(1, 2, 0)  # column halo: box starts at w0-1
    segs = -(-H // seg)
    seg = -(-H // segs)
    for sg in range(segs):
        oh0 = sg * seg
        rows = min(seg, H - oh0)
        for j in range(-1, rows + 1):
            i = oh0 + j
            a_row = xpad[i] if 0 <= i < H else torch.zeros(W + 2, C)  # TMA zero fill outside the image
            k0, k1 = max(0, 1 - j), min(2, rows - j)
            if j + 1 < rows:                                        # first touch of output row j+1
                blk = (g0 + j + 1) % BLOCKS
                assert not touched[blk] and float(tmem[blk].abs().sum()) == 0.0, "accumulator not handed back"
                touched[blk] = True
            if k0 <= k1:
                gfirst = g0 + (j - 1 + k0)
                cnt = k1 - k0 + 1
                b0 = gfirst % BLOCKS
                cnt1 = min(cnt, BLOCKS - b0)
                for s in range(3):
                    a = a_row[s:s + W]                              # descriptor start += s*128 B
                    for part in range(2):
                        c = cnt1 if part == 0 else cnt - cnt1
                        if c <= 0:
                            continue
                        kq = k0 if part == 0 else k0 + cnt1
                        blk = b0 if part == 0 else 0
                        wb = sW[s * 3 + kq:s * 3 + kq + c].reshape(c * 64, 64)   # N = 64*c stacked rows
                        d = a @ wb.t()                              # [W, 64*c]
                        for t in range(c):
                            assert touched[blk + t], "MMA into an accumulator that was not claimed"
                            tmem[blk + t] += d[:, t * 64:(t + 1) * 64]
            if j >= 1:                                              # output row j-1 complete: emit, zero, hand back
                blk = (g0 + j - 1) % BLOCKS
                out[oh0 + j - 1] = tmem[blk]
                tmem[blk].zero_()
                touched[blk] = False
        g0 += rows
    assert not any(touched)
    return out


@pytest.mark.parametrize("H,W,seg", [(1, 9, 16), (2, 12, 16), (5, 7, 2), (9, 10, 4), (23, 6, 16), (16, 5, 5)])
def test_schedule_adds_up_to_the_convolution(H, W, seg):
    g = torch.Generator().manual_seed(H * 100 + W)
    x = torch.randn(H, W, 64, generator=g, dtype=torch.float64).float()
    w = torch.randn(64, 64, 3, 3, generator=g, dtype=torch.float64).float() / 24      # [O][I][R][S]
    # forward: kernel operand w_fwd [O][R][S][I] -> tap tiles [t][n = O][k = I]
    w_fwd = w.permute(0, 2, 3, 1).reshape(64, 9, 64).permute(1, 0, 2).contiguous()
    y = emulate(x, w_fwd, dgrad=False, seg=seg)
    y_ref = F.conv2d(x.permute(2, 0, 1)[None], w, None, 1, 1)[0].permute(1, 2, 0)
    assert float((y - y_ref).abs().max()) <= 1e-4 * max(1.0, float(y_ref.abs().max()))
    # data gradient: operand w_dgrad [I][R][S][O] -> tap tiles [t][n = I][k = O]; input = dy
    dy = torch.randn(H, W, 64, generator=g, dtype=torch.float64).float()
    w_dg = w.permute(1, 2, 3, 0).reshape(64, 9, 64).permute(1, 0, 2).contiguous()
    dx = emulate(dy, w_dg, dgrad=True, seg=seg)
    xin = x.permute(2, 0, 1)[None].clone().requires_grad_(True)
    F.conv2d(xin, w, None, 1, 1).backward(dy.permute(2, 0, 1)[None])
    dx_ref = xin.grad[0].permute(1, 2, 0)
    assert float((dx - dx_ref).abs().max()) <= 1e-4 * max(1.0, float(dx_ref.abs().max()))

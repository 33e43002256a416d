"""Per-kernel parity: every libecgmm kernel family against the plain fp32 PyTorch operator it
replaces, on the same (bf16-rounded where the kernel stores bf16) operands.

Tolerances: kernels that write bf16 are compared with rtol 2^-7 (one bf16 ulp is 2^-8) plus a small
atol scaled to the tensor; fp32-in/fp32-out kernels with 1e-4 relative L2."""
import zlib

import pytest
import torch
import torch.nn.functional as F

import ecgmm  # noqa: F401
from ecgmm import lib, ops
from ecgmm import nn as enn
from ecgmm import optim as eoptim

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF = torch.bfloat16


def gen(name):
    return torch.Generator().manual_seed(zlib.crc32(name.encode()) % (2**31))


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def close_bf16(a, b, what, rtol=2.0**-7, atol_scale=2.0**-7):
    a, b = a.float().cpu(), b.float().cpu()
    atol = atol_scale * float(b.abs().max()) * 0.25 + 1e-6
    bad = (a - b).abs() > atol + rtol * b.abs()
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.numel()} outside tolerance, max err {float((a - b).abs().max()):.4g}"


def nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2).contiguous()


@pytest.fixture(scope="module", autouse=True)
def _device():
    lib.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


# ------------------------------------------------------------------ BatchNorm forward
@pytest.mark.parametrize("N,H,W,C", [(2, 7, 13, 64), (3, 1, 619, 128), (5, 16, 40, 256), (4, 8, 79, 512),
                                     (130, 3, 5, 64)])
@pytest.mark.parametrize("mode", ["plain", "res", "norelu"])
def test_bn_train_forward(N, H, W, C, mode):
    g = gen(f"bnf{N}{H}{W}{C}{mode}")
    x = (torch.randn(N, C, H, W, generator=g) * 1.7 + 0.3).to(DEV).to(BF)
    res = torch.randn(N, C, H, W, generator=g).to(DEV).to(BF) if mode == "res" else None
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.3 * torch.randn(C, generator=g))
        bn.bias.copy_(0.2 * torch.randn(C, generator=g))
        bn.running_mean.copy_(0.1 * torch.randn(C, generator=g))
        bn.running_var.copy_(1 + 0.2 * torch.rand(C, generator=g))
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    bias = (0.5 * torch.randn(C, generator=g)).to(DEV)
    # reference: conv bias added before BN (it cancels in y, shows up in running_mean)
    y_ref = bn(x.float() + bias.view(1, C, 1, 1))
    if res is not None:
        y_ref = y_ref + res.float()
    if mode != "norelu":
        y_ref = F.relu(y_ref)
    rm_ref, rv_ref = bn.running_mean.clone(), bn.running_var.clone()
    xn = nhwc(x)
    st = ops.bn_train_stats(xn, bn.weight.detach(), bn.bias.detach(), rm0, rv0, None, bn.eps, bn.momentum,
                            conv_bias=bias, want_nsum=True)
    y, mask = ops.bn_apply(xn, st, res=None if res is None else nhwc(res), relu=mode != "norelu", want_mask=True)
    close_bf16(nchw(y), y_ref.detach(), "bn forward")
    if mode != "norelu":  # bit j of byte g = (y[8g+j] > 0)
        bits = (y.flatten().view(-1, 8) > 0).to(torch.int32)
        want = (bits << torch.arange(8, device=DEV, dtype=torch.int32)).sum(1)
        assert torch.equal(mask.to(torch.int32), want)
    assert rel_l2(rm0, rm_ref) < 1e-4 and rel_l2(rv0, rv_ref) < 1e-4
    assert rel_l2(st.nsum, x.float().sum(dim=(2, 3))) < 1e-4 or float(st.nsum.abs().max()) < 1e-3


def test_bn_eval_coeffs():
    g = gen("bneval")
    C = 128
    x = torch.randn(2, C, 5, 9, generator=g).to(DEV).to(BF)
    bn = torch.nn.BatchNorm2d(C).to(DEV).eval()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.3 * torch.randn(C, generator=g))
        bn.bias.copy_(0.2 * torch.randn(C, generator=g))
        bn.running_mean.copy_(0.4 * torch.randn(C, generator=g))
        bn.running_var.copy_(0.5 + torch.rand(C, generator=g))
    bias = torch.randn(C, generator=g).to(DEV)
    y_ref = F.relu(bn(x.float() + bias.view(1, C, 1, 1)))
    st = ops.bn_eval_coeffs(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps, bias)
    y, _ = ops.bn_apply(nhwc(x), st, relu=True)
    close_bf16(nchw(y), y_ref, "bn eval")


# ------------------------------------------------------------------ BatchNorm backward
@pytest.mark.parametrize("N,H,W,C", [(2, 7, 13, 64), (3, 1, 310, 128), (4, 8, 79, 512), (70, 4, 6, 256)])
@pytest.mark.parametrize("mode", ["relu", "res_relu", "linear"])
def test_bn_backward(N, H, W, C, mode):
    g = gen(f"bnb{N}{H}{W}{C}{mode}")
    x = (torch.randn(N, C, H, W, generator=g) * 1.3 + 0.2).to(DEV).to(BF)
    res = torch.randn(N, C, H, W, generator=g).to(DEV).to(BF) if mode == "res_relu" else None
    dy = torch.randn(N, C, H, W, generator=g).to(DEV).to(BF)
    gamma = (1 + 0.3 * torch.randn(C, generator=g)).to(DEV).requires_grad_(True)
    beta = (0.2 * torch.randn(C, generator=g)).to(DEV).requires_grad_(True)
    xr = x.float().requires_grad_(True)
    y_ref = F.batch_norm(xr, None, None, gamma, beta, True, 0.1, 1e-5)
    if res is not None:
        y_ref = y_ref + res.float()
    if mode != "linear":
        y_ref = F.relu(y_ref)
    y_ref.backward(dy.float())
    xn = nhwc(x)
    st = ops.bn_train_stats(xn, gamma.detach(), beta.detach(), None, None, None, 1e-5, 0.1)
    y, mask = ops.bn_apply(xn, st, res=None if res is None else nhwc(res), relu=mode != "linear", want_mask=True)
    dg, db = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    dx, dz = ops.bn_backward(xn, nhwc(dy), st, gamma.detach(), y=y if mode != "linear" else None, want_dz=True,
                             dgamma=dg, dbeta=db)
    if mode != "linear":  # the bit-mask path must agree exactly with the y > 0 path
        dg2, db2 = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
        dx2, dz2 = ops.bn_backward(xn, nhwc(dy), st, gamma.detach(), mask=mask, want_dz=True, dgamma=dg2, dbeta=db2)
        assert torch.equal(dx, dx2) and torch.equal(dz, dz2) and torch.equal(dg, dg2) and torch.equal(db, db2)
    # ReLU mask taken from the bf16 output: elements whose pre-activation rounds to 0 may differ -> L2 metric
    assert rel_l2(nchw(dx), xr.grad) < 2e-2
    assert rel_l2(dg, gamma.grad) < 1e-2 and rel_l2(db, beta.grad) < 1e-2
    if mode != "linear":
        mask = (y_ref > 0).float()
        assert rel_l2(nchw(dz), dy.float() * mask) < 1e-2


# ------------------------------------------------------------------ stem BN + ReLU + max-pool
@pytest.mark.parametrize("N,H,W", [(2, 13, 21, ), (3, 1, 75), (2, 32, 50), (1, 125, 64)])
def test_stem_pool_forward_backward(N, H, W):
    C = 64
    g = gen(f"pool{N}{H}{W}")
    x = (torch.randn(N, C, H, W, generator=g) * 1.5).to(DEV).to(BF)
    gamma = (1 + 0.3 * torch.randn(C, generator=g)).to(DEV).requires_grad_(True)
    gamma.data[3] = -0.7  # negative scale: max-pool must run after the affine map
    beta = (0.2 * torch.randn(C, generator=g)).to(DEV).requires_grad_(True)
    xr = x.float().requires_grad_(True)
    a = F.relu(F.batch_norm(xr, None, None, gamma, beta, True, 0.1, 1e-5))
    y_ref = F.max_pool2d(a, 3, 2, 1)
    dyp = torch.randn(y_ref.shape, generator=g).to(DEV).to(BF)
    y_ref.backward(dyp.float())
    xn = nhwc(x)
    st = ops.bn_train_stats(xn, gamma.detach(), beta.detach(), None, None, None, 1e-5, 0.1)
    y, arg = ops.bn_relu_maxpool(xn, st)
    assert tuple(y.shape) == (N, y_ref.shape[2], y_ref.shape[3], C)
    close_bf16(nchw(y), y_ref.detach(), "pool forward")
    dg, db = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    dx, _ = ops.bn_backward(xn, nhwc(dyp), st, gamma.detach(), argmax=arg, dgamma=dg, dbeta=db)
    assert rel_l2(nchw(dx), xr.grad) < 2e-2
    assert rel_l2(dg, gamma.grad) < 1e-2 and rel_l2(db, beta.grad) < 1e-2
    # pooled-domain reduction (mode 4): same sums from the pooled output, no gather
    dg2, db2 = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    dx2, _ = ops.bn_backward(xn, nhwc(dyp), st, gamma.detach(), argmax=arg, dgamma=dg2, dbeta=db2, pooled=y,
                             beta=beta.detach())
    assert rel_l2(nchw(dx2), xr.grad) < 2e-2
    assert rel_l2(dg2, gamma.grad) < 1e-2 and rel_l2(db2, beta.grad) < 1e-2


def test_avgpool():
    g = gen("avg")
    x = torch.randn(5, 8, 79, 512, generator=g).to(DEV).to(BF)
    out = ops.avgpool_fwd(x)
    assert rel_l2(out, x.float().mean(dim=(1, 2))) < 1e-5
    d = torch.randn(5, 512, generator=g).to(DEV)
    dx = ops.avgpool_bwd(d, x.shape)
    close_bf16(dx, (d / (8 * 79)).view(5, 1, 1, 512).expand(5, 8, 79, 512), "avgpool bwd")


# ------------------------------------------------------------------ 1-D stem
@pytest.mark.parametrize("B,Cin,L", [(3, 1, 2476), (2, 12, 5000), (2, 12, 777), (1, 1, 9)])
def test_signal_stem(B, Cin, L):
    g = gen(f"sstem{B}{Cin}{L}")
    x = torch.randn(B, Cin, L, generator=g).to(DEV)
    w = (torch.randn(64, Cin, 7, generator=g) / (7 * Cin) ** 0.5).to(DEV).requires_grad_(True)
    y_ref = F.conv1d(x, w, None, 2, 3)
    y = ops.signal_stem_fwd(x, w.detach())
    assert tuple(y.shape) == (B, 1, y_ref.shape[2], 64)
    close_bf16(y[:, 0].permute(0, 2, 1), y_ref.detach(), "signal stem fwd")
    dy = torch.randn(y_ref.shape, generator=g).to(DEV).to(BF)
    y_ref.backward(dy.float())
    dw = torch.zeros_like(w)
    ops.signal_stem_wgrad(x, dy.permute(0, 2, 1).contiguous().view(B, 1, -1, 64), dw)
    assert rel_l2(dw, w.grad) < 1e-4


# ------------------------------------------------------------------ dense kernels
@pytest.mark.parametrize("M,N,K", [(5, 2, 128), (512, 256, 512), (37, 64, 24), (130, 128, 768),
                                   (4100, 330, 70)])  # the last one is large enough for the 64x64-tile variant
def test_linear(M, N, K):
    g = gen(f"lin{M}{N}{K}")
    x = torch.randn(M, K, generator=g).to(DEV).requires_grad_(True)
    lin = torch.nn.Linear(K, N).to(DEV)
    y_ref = F.relu(lin(x))
    y = ops.linear_fwd(x.detach(), lin.weight.detach(), lin.bias.detach(), relu=True)
    assert rel_l2(y, y_ref) < 1e-5
    dy = torch.randn(M, N, generator=g).to(DEV)
    lin(x).backward(dy)
    dw, db = torch.empty_like(lin.weight), torch.empty_like(lin.bias)
    dx = ops.linear_bwd(x.detach(), lin.weight.detach(), dy, dw=dw, db=db)
    assert rel_l2(dx, x.grad) < 1e-5 and rel_l2(dw, lin.weight.grad) < 1e-5 and rel_l2(db, lin.bias.grad) < 1e-5


@pytest.mark.parametrize("rows,D", [(4, 256), (33, 768), (512, 32)])
def test_layernorm(rows, D):
    g = gen(f"ln{rows}{D}")
    x = (torch.randn(rows, D, generator=g) * 2 + 0.5).to(DEV).requires_grad_(True)
    ln = torch.nn.LayerNorm(D).to(DEV)
    with torch.no_grad():
        ln.weight.copy_(1 + 0.2 * torch.randn(D, generator=g))
        ln.bias.copy_(0.1 * torch.randn(D, generator=g))
    y_ref = ln(x)
    dy = torch.randn(rows, D, generator=g).to(DEV)
    y_ref.backward(dy)
    y, mean, rstd = ops.layernorm_fwd(x.detach(), ln.weight.detach(), ln.bias.detach(), ln.eps)
    assert rel_l2(y, y_ref) < 1e-5
    dg, db = torch.empty(D, device=DEV), torch.empty(D, device=DEV)
    dx = ops.layernorm_bwd(x.detach(), dy, ln.weight.detach(), mean, rstd, dg, db)
    assert rel_l2(dx, x.grad) < 1e-4 and rel_l2(dg, ln.weight.grad) < 1e-4 and rel_l2(db, ln.bias.grad) < 1e-4


@pytest.mark.parametrize("B,C,focal", [(16, 2, 0), (16, 2, 1), (300, 5, 1), (1, 2, 0)])
def test_losses(B, C, focal):
    g = gen(f"loss{B}{C}{focal}")
    z = (torch.randn(B, C, generator=g) * 2).to(DEV).requires_grad_(True)
    y = torch.randint(0, C, (B,), generator=g).to(DEV)
    if focal:
        ce = F.cross_entropy(z, y, reduction="none")
        ref = (1.0 * (1 - torch.exp(-ce)) ** 2.0 * ce).mean()
        crit = enn.FocalLoss()
    else:
        ref = F.cross_entropy(z, y)
        crit = enn.CrossEntropyLoss()
    (ref * 1.7).backward()
    z2 = z.detach().clone().requires_grad_(True)
    out = crit(z2, y)
    (out * 1.7).backward()
    assert abs(float(out) - float(ref)) < 1e-5 * max(1, abs(float(ref)))
    assert rel_l2(z2.grad, z.grad) < 1e-4


def test_zscore_and_dropout():
    g = gen("zs")
    x = (torch.randn(7, 12, 5000, generator=g) * 3 + 1).to(DEV)
    y = ops.zscore(x)
    ref = (x - x.mean(-1, keepdim=True)) / (x.var(-1, unbiased=False, keepdim=True).sqrt() + 1e-8)
    assert rel_l2(y, ref) < 1e-5
    h = torch.ones(1 << 16, device=DEV)
    yd, mask = ops.dropout_fwd(h, 0.3, 1234)
    keep = float((mask > 0).float().mean())
    assert abs(keep - 0.7) < 0.01 and abs(float(yd.mean()) - 1.0) < 0.02
    yd2, _ = ops.dropout_fwd(h, 0.3, 1234)
    assert torch.equal(yd, yd2)
    y0, _ = ops.dropout_fwd(h, 0.0, 1)
    assert torch.equal(y0, h)


@pytest.mark.parametrize("steps", [1, 3])
def test_adam_matches_torch(steps):
    g = gen("adam")
    shapes = [(64, 3, 7, 7), (64,), (3,), (128, 64, 3, 3), (70001,)]
    ps_ref = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    ps = [torch.nn.Parameter(p.detach().clone()) for p in ps_ref]
    o_ref = torch.optim.Adam(ps_ref, lr=1e-3)
    o = eoptim.Adam(ps, lr=1e-3)
    for it in range(steps):
        for a, b in zip(ps_ref, ps):
            gr = torch.randn(a.shape, generator=g).to(DEV) * (0.1 + it)
            a.grad = gr.clone()
            b.grad = gr.clone()
        if it == 2:
            for grp in o_ref.param_groups + o.param_groups:
                grp["lr"] /= 10  # train.py:158-161
        o_ref.step()
        o.step()
    for a, b in zip(ps_ref, ps):
        assert float((a - b).abs().max()) < 2e-6
    sd = o.state_dict()
    assert set(sd["state"][0].keys()) >= {"step", "exp_avg", "exp_avg_sq"}


# ------------------------------------------------------------------ whole 1-D SE block
def test_basic_block_1d_se():
    from ecgmm import model as M
    from oracle.model import BasicBlock1D as RefBlock

    torch.manual_seed(3)
    ref = RefBlock(64, 128, stride=2).to(DEV)
    blk = M.BasicBlock1D(64, 128, stride=2).to(DEV)
    # weights representable in bf16 so that only activation rounding differs
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_(p.to(BF).float())
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(1 + 0.2 * torch.randn_like(m.weight))
                m.bias.copy_(0.1 * torch.randn_like(m.bias))
    blk.load_state_dict(ref.state_dict())
    g = gen("blk1d")
    x = torch.randn(6, 64, 619, generator=g).to(DEV).to(BF)
    xr = x.float().requires_grad_(True)
    out_ref = ref(xr)
    dout = torch.randn(out_ref.shape, generator=g).to(DEV).to(BF)
    out_ref.backward(dout.float())
    xin = x.permute(0, 2, 1).contiguous().view(6, 1, 619, 64)
    out, rec = M._conv_block_fwd(blk, xin, save=True)
    assert rel_l2(out[:, 0].permute(0, 2, 1), out_ref) < 1e-2
    G = M.GradArena(list(blk.parameters()), DEV)
    dx, _ = M._conv_block_bwd(blk, rec, dout.permute(0, 2, 1).contiguous().view(6, 1, -1, 128), G)
    assert rel_l2(dx[:, 0].permute(0, 2, 1), xr.grad) < 5e-2
    scale = max(float(p.grad.norm()) for p in ref.parameters())
    for (k, pr), (_, p) in zip(ref.named_parameters(), blk.named_parameters()):
        if float(pr.grad.norm()) < 1e-4 * scale:
            assert float(G(p).norm()) < 1e-3 * scale, k
            continue
        assert rel_l2(G(p), pr.grad) < 5e-2, (k, rel_l2(G(p), pr.grad))
    for (k, a), (_, b) in zip(ref.state_dict().items(), blk.state_dict().items()):
        if "running" in k:
            assert rel_l2(b, a) < 1e-2, k


@pytest.mark.parametrize("cin,cout,stride,H,W", [(64, 64, 1, 16, 40), (64, 128, 2, 17, 45), (256, 512, 2, 16, 39)])
def test_basic_block_2d(cin, cout, stride, H, W):
    from ecgmm import model as M
    from oracle.model import BasicBlock2D as RefBlock

    torch.manual_seed(5)
    ref = RefBlock(cin, cout, stride).to(DEV)
    blk = M.BasicBlock(cin, cout, stride).to(DEV)
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_(p.to(BF).float())
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(1 + 0.2 * torch.randn_like(m.weight))
                m.bias.copy_(0.1 * torch.randn_like(m.bias))
    blk.load_state_dict(ref.state_dict())
    g = gen(f"blk2d{cin}{cout}")
    N = 4
    x = torch.randn(N, cin, H, W, generator=g).to(DEV).to(BF)
    xr = x.float().requires_grad_(True)
    out_ref = ref(xr)
    dout = torch.randn(out_ref.shape, generator=g).to(DEV).to(BF)
    out_ref.backward(dout.float())
    out, rec = M._conv_block_fwd(blk, nhwc(x), save=True)
    assert rel_l2(nchw(out), out_ref) < 1e-2
    G = M.GradArena(list(blk.parameters()), DEV)
    dx, _ = M._conv_block_bwd(blk, rec, nhwc(dout), G)
    assert rel_l2(nchw(dx), xr.grad) < 5e-2
    for (k, pr), (_, p) in zip(ref.named_parameters(), blk.named_parameters()):
        assert rel_l2(G(p), pr.grad) < 5e-2, (k, rel_l2(G(p), pr.grad))


# ------------------------------------------------------------------ uint8 image input
def test_stem_s2d_uint8_equals_totensor_normalize():
    """Raw uint8 pixels through the stem staging kernel == torchvision ToTensor + Normalize(0.5, 0.5)
    (dataset.py:119-123) applied on the host and fed as fp32: bit-exact, every one of the 256 levels."""
    g = gen("u8")
    u8 = torch.randint(0, 256, (3, 3, 37, 91), generator=g, dtype=torch.uint8)
    u8[0, :, 0, :86] = torch.arange(256, dtype=torch.uint8)[:86]
    u8[1].view(-1)[:256] = torch.arange(256, dtype=torch.uint8)
    ref = ((u8.float() / 255.0) - 0.5) / 0.5
    a = ops.stem_s2d(u8.to(DEV))
    b = ops.stem_s2d(ref.to(DEV))
    assert torch.equal(a, b)
    with pytest.raises(lib.EcgmmError):
        ops.stem_s2d(u8.to(DEV).to(torch.int32))


# ------------------------------------------------------------------ BatchNorm statistics from the conv epilogues
@pytest.mark.parametrize("N,H,W,Cin,Cout,k,stride", [
    (3, 21, 150, 64, 64, 3, 1),     # 64 -> 64 rolling-accumulator kernel: sums kept in registers, one row per CTA
    (4, 63, 625, 64, 64, 3, 1),     # ... at the native layer1 size (many tiles per CTA)
    (2, 1, 300, 64, 64, 3, 1),      # 1-D 64 -> 64 (halo kernel): not offered
    (2, 9, 40, 128, 64, 3, 1),      # generic kernel, BN = 64
    (3, 17, 45, 64, 128, 3, 2),     # stride 2 with K = 576: below the K >= 1152 threshold, not offered
    (3, 17, 45, 128, 128, 3, 2),    # stride 2, BN = 128
    (2, 12, 31, 128, 256, 1, 2),    # 1x1 downsample, BN = 256
    (2, 9, 20, 128, 256, 3, 1),     # BN = 256, ragged tiles
    (2, 8, 79, 256, 512, 3, 1),     # two N tiles per pixel tile
    (5, 1, 310, 512, 128, 3, 1),    # 1-D (H = 1), K = 1536
])
@pytest.mark.parametrize("pair", ["1", "0"], ids=["cta_pair", "single_cta"])
def test_conv_epilogue_statistics(N, H, W, Cin, Cout, k, stride, pair, monkeypatch):
    """conv2d_fwd(want_stats=True): same y as without, and the partial rows fold to the exact per-channel sum /
    sum of squares of the STORED bf16 tensor (fp32 partials over <= a few thousand values each)."""
    monkeypatch.setenv("ECGMM_NT_PAIR", pair)
    ops._SHAPE_CACHE.clear()  # the number of partial rows depends on the kernel
    g = gen(f"cs{N}{H}{W}{Cin}{Cout}{k}{stride}")
    x = torch.randn(N, H, W, Cin, generator=g).to(DEV).to(torch.bfloat16)
    R = 1 if H == 1 else k
    w = (torch.randn(Cout, Cin, R, k, generator=g) / (Cin * R * k) ** 0.5).to(DEV)
    w_fwd, _ = ops.conv_weight_prep(w)
    y0 = ops.conv2d_fwd(x, w_fwd, stride)
    y, part = ops.conv2d_fwd(x, w_fwd, stride, want_stats=True)
    assert torch.equal(y, y0)
    stack = (Cin == 64 and Cout == 64 and k == 3 and H > 1 and stride == 1 and W >= 96)
    if k * (1 if H == 1 else k) * Cin < 1152 and not stack:
        assert part is None  # not offered (too few MMAs per tile to hide it): the caller runs the statistics pass
        return
    yd = y.double().reshape(-1, Cout)
    s_ref, q_ref = yd.sum(0), (yd * yd).sum(0)
    s = part.psum.view(part.rows, Cout).double().sum(0)
    q = part.psq.view(part.rows, Cout).double().sum(0)
    scale = yd.abs().sum(0)
    assert ((s - s_ref).abs() <= 1e-5 * scale + 1e-6).all()
    assert ((q - q_ref).abs() <= 1e-5 * q_ref + 1e-6).all()
    # and through the finalize kernel: identical BatchNorm coefficients to the separate statistics pass
    bn = torch.nn.BatchNorm2d(Cout).to(DEV)
    a = ops.bn_train_stats(y, bn.weight, bn.bias, None, None, None, 1e-5, 0.1)
    b = ops.bn_train_stats(y, bn.weight, bn.bias, None, None, None, 1e-5, 0.1, partials=part)
    assert rel_l2(b.mean, a.mean) < 1e-4 and rel_l2(b.invstd, a.invstd) < 1e-5


@pytest.mark.parametrize("N,H,W", [(2, 50, 100), (3, 37, 75), (2, 250, 2500)])
def test_stem_epilogue_statistics(N, H, W):
    """stem_conv_fwd(want_stats=True): same y, and the per-CTA partial rows (sums kept in registers for the whole
    kernel) fold to the per-channel sum / sum of squares of the STORED bf16 tensor."""
    g = gen(f"stemstats{N}{H}{W}")
    image = torch.randn(N, 3, H, W, generator=g).clamp(-1, 1).to(DEV)
    w = (torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5).to(DEV)
    xs, ws = ops.stem_s2d(image), ops.stem_weight_prep(w)
    y0 = ops.stem_conv_fwd(xs, ws, H, W)
    y, part = ops.stem_conv_fwd(xs, ws, H, W, want_stats=True)
    assert part is not None and torch.equal(y, y0)
    yd = y.double().reshape(-1, 64)
    s = part.psum.view(part.rows, 64).double().sum(0)
    q = part.psq.view(part.rows, 64).double().sum(0)
    assert ((s - yd.sum(0)).abs() <= 1e-5 * yd.abs().sum(0) + 1e-6).all()
    assert ((q - (yd * yd).sum(0)).abs() <= 1e-5 * (yd * yd).sum(0) + 1e-6).all()


@pytest.mark.parametrize("N,H,W,accumulate,with_mask", [(3, 21, 150, False, True), (2, 63, 625, True, True),
                                                         (2, 5, 130, True, False)])
def test_dgrad_epilogue_reduces_for_batchnorm_backward(N, H, W, accumulate, with_mask):
    """conv2d_dgrad(reduce_for=(x, mask, stats)): same dx as without, and the partial rows fold to sum dz and
    sum dz * xhat of the STORED dx against that BatchNorm's input -- so bn_backward(partials=...) returns what the
    separate reduction pass returns."""
    g = gen(f"dgred{N}{H}{W}{accumulate}{with_mask}")
    C = 64
    dy = torch.randn(N, H, W, C, generator=g).to(DEV).to(BF)
    w = (torch.randn(C, C, 3, 3, generator=g) / 24).to(DEV)
    _, w_dg = ops.conv_weight_prep(w)
    xbn = (torch.randn(N, H, W, C, generator=g) * 1.5 + 0.3).to(DEV).to(BF)
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    st = ops.bn_train_stats(xbn, bn.weight, bn.bias, None, None, None, 1e-5, 0.1)
    mask = None
    if with_mask:
        _, mask = ops.bn_apply(xbn, st, relu=True, want_mask=True)
    old = torch.randn(N, H, W, C, generator=g).to(DEV).to(BF) if accumulate else None
    ref_dx = ops.conv2d_dgrad(dy, w_dg, (H, W), 1, out=None if old is None else old.clone(), accumulate=accumulate)
    dx, part = ops.conv2d_dgrad(dy, w_dg, (H, W), 1, out=None if old is None else old.clone(), accumulate=accumulate,
                                reduce_for=(xbn, mask, st))
    assert part is not None and torch.equal(dx, ref_dx)
    dz = dx.double()
    if with_mask:
        bits = torch.stack([(mask.view(N, H, W, C // 8) >> j) & 1 for j in range(8)], dim=-1).reshape(N, H, W, C)
        dz = dz * bits.double()
    xhat = (xbn.double() - st.mean.double()) * st.invstd.double()
    s1, s2 = dz.reshape(-1, C).sum(0), (dz * xhat).reshape(-1, C).sum(0)
    p1 = part.p1.view(part.rows, C).double().sum(0)
    p2 = part.p2.view(part.rows, C).double().sum(0)
    scale1, scale2 = dz.abs().reshape(-1, C).sum(0), (dz * xhat).abs().reshape(-1, C).sum(0)
    assert ((p1 - s1).abs() <= 2e-5 * scale1 + 1e-5).all()
    assert ((p2 - s2).abs() <= 2e-5 * scale2 + 1e-5).all()
    dg_a, db_a, dg_b, db_b = (torch.zeros(C, device=DEV) for _ in range(4))
    a, _ = ops.bn_backward(xbn, dx, st, bn.weight, mask=mask, dgamma=dg_a, dbeta=db_a)
    b, _ = ops.bn_backward(xbn, dx, st, bn.weight, mask=mask, dgamma=dg_b, dbeta=db_b, partials=part)
    assert rel_l2(b, a) < 2e-3 and rel_l2(dg_b, dg_a) < 1e-4 and rel_l2(db_b, db_a) < 1e-4


@pytest.mark.parametrize("B,Cin,L", [(4, 12, 5000), (3, 1, 2476), (2, 12, 603), (2, 5, 64)])
def test_signal_stem_on_tensor_cores(B, Cin, L):
    """Conv1d(Cin, 64, k 7, stride 2, pad 3) as a 1x3 convolution over 4-sample groups (signal_s4d + the generic tcgen05
    forward / weight-gradient kernels) against torch's conv1d on the same bf16-rounded operands."""
    g = gen(f"s4d{B}{Cin}{L}")
    x = torch.randn(B, Cin, L, generator=g).to(DEV)
    w = (torch.randn(64, Cin, 7, generator=g) / (7 * Cin) ** 0.5).to(DEV)
    xs4 = ops.signal_s4d(x)
    assert xs4 is not None
    y = ops.conv2d_fwd(xs4, ops.signal_stem_w4(w), 1).view(B, -1, 64)
    wr = w.to(BF).float().requires_grad_(True)
    ref = torch.nn.functional.conv1d(x.to(BF).float(), wr, None, 2, 3)
    assert y.shape[1] == ref.shape[2]
    assert rel_l2(y.float().permute(0, 2, 1), ref) < 6e-3
    dy = torch.randn(ref.shape, generator=g).to(DEV).to(BF)
    ref.backward(dy.float())
    dw = torch.zeros_like(w)
    ops.signal_stem_wgrad_s4d(xs4, dy.permute(0, 2, 1).contiguous().view(B, 1, -1, 64), dw)
    assert rel_l2(dw, wr.grad) < 2e-3
    assert ops.signal_s4d(torch.zeros(2, Cin, 601, device=DEV)) is None  # L mod 4 == 1: direct kernels


def test_batched_weight_prep_is_bit_identical_to_the_single_tensor_kernel():
    """ecgmm_conv_weight_prep_batch (all stale convolution weights of a stage in one launch) against
    ecgmm_conv_weight_prep, for every weight shape of the two ResNets plus a channel count it hands back (12)."""
    from ecgmm import ops

    g = torch.Generator().manual_seed(3)
    shapes = [(64, 64, 3, 3), (128, 64, 3, 3), (128, 64, 1, 1), (128, 128, 3, 3), (256, 128, 3, 3), (256, 128, 1, 1),
              (256, 256, 3, 3), (512, 256, 3, 3), (512, 256, 1, 1), (512, 512, 3, 3), (64, 64, 3), (128, 64, 3),
              (128, 64, 1), (256, 256, 3), (64, 12, 7), (96, 32, 2, 2)]
    ws = [torch.randn(*sh, generator=g).cuda() for sh in shapes]
    got = ops.conv_weight_prep_batch(ws * 3)  # 48 tensors: more than one launch of 32
    for k, (f, d) in enumerate(got):
        rf, rd = ops.conv_weight_prep(ws[k % len(ws)])
        assert torch.equal(f, rf) and torch.equal(d, rd), shapes[k % len(ws)]


@pytest.mark.parametrize("N,H,W", [(2, 125, 1250), (3, 33, 70), (1, 8, 32), (2, 9, 35), (1, 40, 33)])
def test_tiled_stem_maxpool_is_bit_identical(N, H, W, monkeypatch):
    """bn_relu_maxpool_tiled_kernel (window staged in shared memory, C = 64) against the per-output kernel: outputs and
    argmax codes must be identical, including ragged tiles and the image border."""
    from ecgmm import ops

    g = gen(f"tiledpool{N}{H}{W}")
    x = nhwc((torch.randn(N, 64, H, W, generator=g) * 1.5).to(DEV).to(BF))
    gamma = (1 + 0.3 * torch.randn(64, generator=g)).to(DEV)
    gamma[5] = -0.6
    beta = (0.2 * torch.randn(64, generator=g)).to(DEV)
    st = ops.bn_train_stats(x, gamma, beta, None, None, None, 1e-5, 0.1)
    y1, a1 = ops.bn_relu_maxpool(x, st)  # default: branch-free scan over a flipped-domain tile
    monkeypatch.setenv("ECGMM_POOL_TILED_V1", "1")
    y2, a2 = ops.bn_relu_maxpool(x, st)  # first tiled version (bounds tests in the scan)
    monkeypatch.setenv("ECGMM_POOL_LEGACY", "1")
    y0, a0 = ops.bn_relu_maxpool(x, st)
    assert torch.equal(y0, y1) and torch.equal(a0, a1)
    assert torch.equal(y0, y2) and torch.equal(a0, a2)


@pytest.mark.parametrize("N,H,W", [(2, 125, 1250), (3, 33, 70), (2, 9, 35), (1, 40, 33), (5, 1, 75)])
def test_stem_bwd_apply_variants_are_bit_identical(N, H, W, monkeypatch):
    """stem_bwd_apply_rows_kernel<1|2> (one CTA per pooled row; coefficients in registers / shared memory) against the
    grid-stride kernel: same arithmetic per element, so dx must be identical, odd heights / widths included."""
    from ecgmm import ops

    g = gen(f"stembwd{N}{H}{W}")
    x = nhwc((torch.randn(N, 64, H, W, generator=g) * 1.5).to(DEV).to(BF))
    gamma = (1 + 0.3 * torch.randn(64, generator=g)).to(DEV)
    gamma[9] = -0.4
    beta = (0.2 * torch.randn(64, generator=g)).to(DEV)
    st = ops.bn_train_stats(x, gamma, beta, None, None, None, 1e-5, 0.1)
    y, arg = ops.bn_relu_maxpool(x, st)
    dyp = torch.randn(y.shape, generator=g).to(DEV).to(BF)
    outs = []
    for v in ("0", "1", "2"):
        monkeypatch.setenv("ECGMM_STEM_BWD_APPLY", v)
        dg, db = torch.empty(64, device=DEV), torch.empty(64, device=DEV)
        dx, _ = ops.bn_backward(x, dyp, st, gamma, argmax=arg, dgamma=dg, dbeta=db, pooled=y, beta=beta)
        outs.append((dx, dg, db))
    for o in outs[1:]:
        assert all(torch.equal(p, q) for p, q in zip(o, outs[0]))


@pytest.mark.parametrize("N,H,W,C", [(3, 63, 625, 64), (2, 7, 13, 64), (5, 1, 310, 128), (4, 8, 79, 512), (3, 5, 9, 2048),
                                     (1, 125, 1250, 64), (70, 16, 157, 256), (2, 1, 1, 64)])
def test_bn_fast_paths_are_bit_identical(N, H, W, C, monkeypatch):
    """The BatchNorm fast paths against the generic kernels: (ECGMM_BN_FAST, ECGMM_BN_ASYNC) = (0, 0) generic apply /
    backward-apply and register reduction; (1, 0) coefficients in registers, ReLU bit mask applied to the packed gradient
    words; (1, 1) the default: backward-apply and backward-reduce with their loads staged through a thread-private
    cp.async ring.  Same arithmetic per element, same summation order -> identical bits."""
    from ecgmm import ops

    g = gen(f"bnfast{N}{H}{W}{C}")
    x = nhwc((torch.randn(N, C, H, W, generator=g) * 1.3 + 0.2).to(DEV).to(BF))
    res = nhwc(torch.randn(N, C, H, W, generator=g).to(DEV).to(BF))
    dy = nhwc(torch.randn(N, C, H, W, generator=g).to(DEV).to(BF))
    dy.view(-1)[::7] = -0.0  # negative zeros must survive the packed masking as they survive unpack -> pack
    gamma = (1 + 0.3 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(C, generator=g)).to(DEV)
    st = ops.bn_train_stats(x, gamma, beta, None, None, None, 1e-5, 0.1)
    outs = []
    for fast, asyn in (("0", "0"), ("1", "0"), ("1", "1")):
        monkeypatch.setenv("ECGMM_BN_FAST", fast)
        monkeypatch.setenv("ECGMM_BN_ASYNC", asyn)
        o = []
        for kw in (dict(relu=True, want_mask=True), dict(res=res, relu=True, want_mask=True), dict(res=res, relu=False),
                   dict(relu=False)):
            y, m = ops.bn_apply(x, st, **kw)
            o += [y] + ([m] if m is not None else [])
        mask = o[3]  # of the residual + ReLU variant
        for kw in (dict(mask=mask, want_dz=True), dict(mask=mask), dict(want_dz=True), dict()):
            dg, db = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
            dx, dz = ops.bn_backward(x, dy, st, gamma, dgamma=dg, dbeta=db, **kw)
            o += [dx, dg, db] + ([dz] if dz is not None else [])
        outs.append(o)
    bits = lambda t: t.view(torch.int16) if t.dtype == BF else t  # noqa: E731  (compare -0.0 / +0.0 as bit patterns)
    for v, o in enumerate(outs[1:], 1):
        assert len(o) == len(outs[0])
        for k, (p, q) in enumerate(zip(outs[0], o)):
            assert torch.equal(bits(p), bits(q)), (v, k)

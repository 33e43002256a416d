"""tcgen05 implicit-GEMM convolutions (forward / data-gradient / weight-gradient / stem) against
torch's fp32 convolution on the same bf16-rounded operands (tests/gpu_conv_check.py)."""
import os

import pytest
import torch

import gpu_conv_check as chk
from ecgmm import lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device():
    lib.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("case", chk.CASES, ids=[c[0] for c in chk.CASES])
def test_conv_case(case):
    results = []
    assert chk.run_case(*case, results=results)
    bad = [r for r in results if not r[2]]
    assert not bad, bad


@pytest.mark.parametrize("bn", ["64", "128"])
@pytest.mark.parametrize("case", [c for c in chk.CASES if c[7] == 3 and c[8] == 1 and c[5] % 128 == 0],
                         ids=lambda c: c[0] if isinstance(c, tuple) else str(c))
def test_wgrad_both_mma_widths(case, bn, monkeypatch):
    """The weight-gradient kernel picks N = 64 or N = 128 MMAs (two CTA types) per layer; force each on every
    stride-1 layer with Cout % 128 == 0 (the library reads ECGMM_WG_BN at every call)."""
    from ecgmm import ops

    monkeypatch.setenv("ECGMM_WG_BN", bn)
    monkeypatch.setenv("ECGMM_WG_T", "0")  # the default for these layers is the transposed kernel (next test)
    ops._SHAPE_CACHE.clear()  # workspace sizes depend on the shape
    results = []
    assert chk.run_case(*case, results=results)
    bad = [r for r in results if not r[2]]
    assert not bad, bad
    ops._SHAPE_CACHE.clear()


@pytest.mark.parametrize("case", [c for c in chk.CASES if c[6] == 3 and c[7] == 3 and c[8] == 1 and c[5] % 128 == 0],
                         ids=lambda c: c[0] if isinstance(c, tuple) else str(c))
def test_transposed_wgrad(case, monkeypatch):
    """The default for 3x3 layers with Cout % 128 == 0: M = 128 output channels x N = 192 (three horizontal taps x 64
    input channels) MMAs, two CTA types (wgrad_halo_kernel<128, true>); bookkeeping emulated in
    tests/test_wgrad_t_emulation_cpu.py.  (test_layer_case runs it too; this test pins the switch explicitly.)"""
    from ecgmm import ops

    monkeypatch.setenv("ECGMM_WG_T", "1")
    ops._SHAPE_CACHE.clear()  # workspace sizes depend on the mode
    results = []
    assert chk.run_case(*case, results=results, do=("wgrad",))
    bad = [r for r in results if not r[2]]
    assert not bad, bad
    ops._SHAPE_CACHE.clear()


PAIR_CASES = [c for c in chk.CASES if c[4] % 128 == 0 or c[5] % 128 == 0] + [
    ("pair_odd_tiles_128", 1, 9, 40, 128, 128, 3, 3, 1),      # 3 pixel tiles: the last pair has a dummy second tile
    ("pair_one_tile_256", 1, 4, 20, 256, 256, 3, 3, 1),       # a single tile: falls back to the single-CTA kernel
    ("pair_layer3_16x157", 4, 16, 157, 256, 256, 3, 3, 1),
    ("pair_layer4_8x79", 6, 8, 79, 512, 512, 3, 3, 1),
]


@pytest.mark.parametrize("pair", ["1", "0"], ids=["cta_pair", "single_cta"])
@pytest.mark.parametrize("case", PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
def test_cta_pair_kernels(case, pair, monkeypatch):
    """Forward / data-gradient (+ accumulating) of every shape whose GEMM N is a multiple of 128 through
    igemm_nt_pair_kernel (tcgen05 cta_group::2: two CTAs, one M = 256 MMA, half of B per CTA) and through the
    single-CTA igemm_nt_kernel."""
    monkeypatch.setenv("ECGMM_NT_PAIR", pair)
    results = []
    assert chk.run_case(*case, results=results, do=("fwd", "dgrad"))
    bad = [r for r in results if not r[2]]
    assert not bad, bad


@pytest.mark.parametrize("name,N,H,W", [("stem_small", 2, 50, 100), ("stem_odd", 1, 37, 75),
                                        ("stem_250x2500", 2, 250, 2500)])
def test_stem_case(name, N, H, W):
    results = []
    assert chk.run_stem(name, N, H, W, results)
    bad = [r for r in results if not r[2]]
    assert not bad, bad


def test_conv_is_linear_at_full_size():
    """Size-independent property at the native layer1 shape: conv(a) + conv(b) == conv(a + b)
    up to bf16 rounding, and an all-zero input gives an all-zero output."""
    from ecgmm import ops

    g = torch.Generator().manual_seed(0)
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24).cuda()
    w_fwd, _ = ops.conv_weight_prep(w)
    a = torch.randn(2, 63, 625, 64, generator=g).cuda().to(torch.bfloat16)
    b = torch.randn(2, 63, 625, 64, generator=g).cuda().to(torch.bfloat16)
    s = (a.float() + b.float()).to(torch.bfloat16)
    ya, yb, ys = ops.conv2d_fwd(a, w_fwd), ops.conv2d_fwd(b, w_fwd), ops.conv2d_fwd(s, w_fwd)
    err = (ya.float() + yb.float() - ys.float()).abs().max().item()
    assert err < 0.06, err
    z = ops.conv2d_fwd(torch.zeros_like(a), w_fwd)
    assert float(z.float().abs().max()) == 0.0


def test_transposed_wgrad_coalesced_fold_is_bit_identical(monkeypatch):
    """wgrad_halo_reduce_t3_kernel (coalesced writes through a shared-memory tile) against the one-thread-per-element
    fold it replaces: same summation order per element, so the weight gradients are equal bit for bit."""
    from ecgmm import ops

    g = torch.Generator().manual_seed(5)
    for N, H, W, Cin, Cout in ((3, 16, 157, 256, 256), (2, 9, 33, 192, 256), (4, 8, 79, 512, 512)):
        x = torch.randn(N, H, W, Cin, generator=g).cuda().to(torch.bfloat16)
        dy = torch.randn(N, H, W, Cout, generator=g).cuda().to(torch.bfloat16)
        out = []
        for legacy in ("1", None):
            monkeypatch.setenv("ECGMM_WG_REDUCE_T_LEGACY", legacy or "0")
            dw = torch.zeros(Cout, Cin, 3, 3, device="cuda")
            ops.conv2d_wgrad(x, dy, dw, 3, 3, 1)
            out.append(dw)
        assert torch.equal(out[0], out[1]), (Cin, Cout)


def test_generic_wgrad_tiled_fold_is_bit_identical(monkeypatch):
    """tn_reduce_tile_kernel (stride-2 3x3 and 1x1 weight gradients) against the one-thread-per-element fold."""
    from ecgmm import ops

    g = torch.Generator().manual_seed(6)
    for N, H, W, Cin, Cout, R, stride in ((3, 17, 45, 64, 128, 3, 2), (2, 16, 39, 256, 512, 3, 2), (3, 17, 45, 64, 128, 1, 2),
                                          (2, 8, 16, 64, 64, 1, 1), (2, 16, 40, 128, 256, 1, 2)):
        Ho, Wo = (H + 2 * (R // 2) - R) // stride + 1, (W + 2 * (R // 2) - R) // stride + 1
        x = torch.randn(N, H, W, Cin, generator=g).cuda().to(torch.bfloat16)
        dy = torch.randn(N, Ho, Wo, Cout, generator=g).cuda().to(torch.bfloat16)
        out = []
        for legacy in ("1", None):
            monkeypatch.setenv("ECGMM_TN_REDUCE_LEGACY", legacy or "0")
            dw = torch.zeros(Cout, Cin, R, R, device="cuda")
            ops.conv2d_wgrad(x, dy, dw, R, R, stride)
            out.append(dw)
        assert torch.equal(out[0], out[1]), (Cin, Cout, R, stride)


@pytest.mark.parametrize("stack", ["1", "0"], ids=["rolling_accumulators", "halo_kernel"])
@pytest.mark.parametrize("case", [("stack_63x625", 2, 63, 625, 64, 64, 3, 3, 1), ("stack_ragged", 3, 21, 150, 64, 64, 3, 3, 1),
                                  ("stack_one_row", 2, 1, 200, 64, 64, 3, 3, 1), ("stack_two_rows", 1, 2, 130, 64, 64, 3, 3, 1)],
                         ids=lambda c: c[0])
def test_layer1_kernels(case, stack, monkeypatch):
    """Forward, data-gradient and accumulating data-gradient of the 64 -> 64 3x3 layers through igemm_nt_stack_kernel
    (the default: N = 192 MMAs into a ring of output-row accumulators in TMEM) and through igemm_nt_halo_kernel
    (ECGMM_NT_STACK=0)."""
    monkeypatch.setenv("ECGMM_NT_STACK", stack)
    results = []
    assert chk.run_case(*case, results=results, do=("fwd", "dgrad"))
    bad = [r for r in results if not r[2]]
    assert not bad, bad

"""Drop-in ECGMultimodalModel whose forward/backward run on libecgmm (sm_100a) kernels.

Mirrors the module tree, constructor, attribute names, 6-tuple output and 229-key state_dict of
the reference fusion model (G2: /root/reference/multimodal_paper_modal_balance.py:197-354; the
`dims` option covers the 512/128/32 layout of multimodal.py:333-469 minus its TabNet encoder).

torch is only the host here: nn.Module / nn.Parameter hold the fp32 master weights (so
state_dict(), .to(), optimizers and checkpoints behave exactly like the reference) and
torch.autograd connects four hand-written stages -- image encoder, signal encoder, clinical
encoder, fusion head -- each of which is ONE autograd node whose forward and backward are
explicit sequences of C-ABI calls (ecgmm.ops).  No torch operator computes anything on the hot
path and nothing falls back to ATen: leaf containers (Conv2d, BatchNorm2d, ...) refuse to be
called directly.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import lib, ops

F32 = torch.float32
BF16 = torch.bfloat16


# ============================================================================ helpers
class ArenaLayout:
    """Offsets of a stage's parameters inside its gradient arena (REVERSE execution order; every slice
    16-byte aligned).  Built once per stage and reused by every backward (see _Stage.new_arena)."""

    def __init__(self, params_exec_order):
        self.offsets = {}
        off = 0
        for p in reversed(list(params_exec_order)):
            if id(p) in self.offsets:
                continue
            self.offsets[id(p)] = (off, p.numel(), p.shape)
            off += (p.numel() + 3) // 4 * 4
        self.total = off


class GradArena:
    """One flat, zero-initialised fp32 buffer holding the gradients of a stage's parameters.

    Parameters are laid out in REVERSE execution order, so the slices that backward completes
    first are contiguous at the front: the data-parallel wrapper all-reduces prefix ranges
    (buckets) while the rest of backward is still running."""

    def __init__(self, params_exec_order, device):
        lay = params_exec_order if isinstance(params_exec_order, ArenaLayout) else ArenaLayout(params_exec_order)
        self.offsets = lay.offsets
        self.total = lay.total
        self.flat = torch.zeros(lay.total, dtype=F32, device=device)

    def __call__(self, p):
        # A FRESH view every time, never cached: autograd's AccumulateGrad adopts the returned tensor as p.grad only
        # when nobody else holds a reference to it; otherwise it clones it on the spot -- i.e. BEFORE the
        # data-parallel all-reduce, which runs asynchronously on this very memory, has delivered the averaged values.
        off, n, shape = self.offsets[id(p)]
        return self.flat[off:off + n].view(shape)

    def end_of(self, p):
        off, n, _ = self.offsets[id(p)]
        return (off + n + 3) // 4 * 4


def _dropout_seed():
    # CPU generator: reproducible under torch.manual_seed, no device synchronisation
    return int(torch.randint(0, 2**62, (1,)).item())


class _ContainerMixin:
    """Leaf modules are parameter containers; compute happens in the stage that owns them."""

    def forward(self, *a, **k):  # noqa: D401
        raise lib.EcgmmError(
            f"{type(self).__name__} is a parameter container of the fused ecgmm stage that owns it; "
            "call the encoder / head module instead (there is no per-layer ATen fallback)")


class _ShadowMixin:
    """bf16 implicit-GEMM operand copies of an fp32 master weight, refreshed when it changes."""

    def shadows(self):
        w = self.weight
        key = (w.data_ptr(), w._version)
        if getattr(self, "_shadow_key", None) != key:
            self._shadow = self._make_shadows(w.detach())
            self._shadow_key = key
        return self._shadow

    def _make_shadows(self, w):
        return ops.conv_weight_prep(w, need_dgrad=True)


def _refresh_conv_shadows(stage):
    """Refresh the stale bf16 shadows of all plain convolutions of a stage in ONE launch (after an optimizer step every
    weight is stale: 27 single-tensor launches per training step otherwise)."""
    convs = stage.__dict__.get("_plain_convs")
    if convs is None:
        convs = stage.__dict__["_plain_convs"] = [m for m in stage.modules() if type(m) in (Conv2d, Conv1d)]
    stale = [m for m in convs if getattr(m, "_shadow_key", None) != (m.weight.data_ptr(), m.weight._version)]
    if len(stale) < 2:
        return
    for m, sh in zip(stale, ops.conv_weight_prep_batch([m.weight.detach() for m in stale])):
        m._shadow = sh
        m._shadow_key = (m.weight.data_ptr(), m.weight._version)


class Conv2d(_ShadowMixin, _ContainerMixin, nn.Conv2d):
    pass


class Conv1d(_ShadowMixin, _ContainerMixin, nn.Conv1d):
    pass


class StemConv2d(_ShadowMixin, _ContainerMixin, nn.Conv2d):
    def _make_shadows(self, w):
        return (ops.stem_weight_prep(w), None)


class SignalStemConv1d(_ShadowMixin, _ContainerMixin, nn.Conv1d):
    def _make_shadows(self, w):
        return (ops.signal_stem_w4(w), None)


class BatchNorm2d(_ContainerMixin, nn.BatchNorm2d):
    pass


class BatchNorm1d(_ContainerMixin, nn.BatchNorm1d):
    pass


class _Marker(_ContainerMixin, nn.Module):
    """Stateless placeholder keeping the reference's child names / Sequential indices."""


class ReLU(_Marker):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace


class MaxPool(_Marker):
    pass


class AvgPool(_Marker):
    pass


class Flatten(_Marker):
    pass


class Sigmoid(_Marker):
    pass


FUSED_STATS = os.environ.get("ECGMM_FUSED_STATS", "1") != "0"  # BatchNorm sums from the convolution epilogue
SIGNAL_STEM_TC = os.environ.get("ECGMM_SIGNAL_STEM_TC", "1") != "0"  # Conv1d(Cin,64,7,2) stem on the tensor cores


def _conv_bn(conv_fn, bn, *args, want_nsum=False):
    """conv -> (y, BatchNorm statistics).  In training mode the convolution epilogue emits the per-channel partial
    sums (no statistics pass over y) unless per-sample sums are wanted too (SE squeeze)."""
    if bn.training and FUSED_STATS and not want_nsum:
        y, part = conv_fn(*args, want_stats=True)
        return y, part
    return conv_fn(*args), None


def _bn_stats(bn, x, conv_bias=None, want_nsum=False, partials=None):
    """Training: batch statistics (+ running update); eval: running statistics."""
    if bn.training:
        if bn.momentum is None:
            raise lib.EcgmmError("BatchNorm momentum=None (cumulative average) is not supported")
        return ops.bn_train_stats(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                  bn.eps, bn.momentum, conv_bias, want_nsum, partials)
    st = ops.bn_eval_coeffs(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, conv_bias)
    if want_nsum:  # SE squeeze needs the per-sample sums even when the statistics are frozen
        tmp = ops.bn_train_stats(x, None, None, None, None, None, bn.eps, 0.0, None, True)
        st.nsum = tmp.nsum
    return st


# ============================================================================ autograd nodes
class _StageFn(torch.autograd.Function):
    """Generic bridge: stage.run_forward(*inputs) -> (outputs, state); stage.run_backward(state, *grads)."""

    @staticmethod
    def forward(ctx, stage, n_inputs, *tensors):
        inputs, params = tensors[:n_inputs], tensors[n_inputs:]
        save = any(ctx.needs_input_grad[2:])
        pre = stage.__dict__.pop("_precomputed", None)
        if pre is not None and pre[0] == (save, tuple((t.data_ptr(), t._version) for t in inputs)):
            outs, state = pre[1]  # kernels already enqueued by _Stage.precompute(); this call only creates the node
        else:
            outs, state = stage.run_forward(*inputs, save=save)
        ctx.stage, ctx.state, ctx.n_inputs, ctx.n_params = stage, state, n_inputs, len(params)
        ctx.backward_ok = stage.training or getattr(stage, "eval_backward_ok", False)
        single = not isinstance(outs, tuple)
        ctx.single = single
        nd = getattr(stage, "non_differentiable_outputs", ())
        if not single and nd:
            ctx.mark_non_differentiable(*[outs[i] for i in nd])
        return outs

    @staticmethod
    def backward(ctx, *grads):
        if ctx.state is None:
            raise lib.EcgmmError("backward through an ecgmm stage whose forward did not record state")
        if not ctx.backward_ok:
            raise lib.EcgmmError(
                f"backward through {type(ctx.stage).__name__} in eval mode is not implemented "
                "(BatchNorm backward uses batch statistics); call .train() first")
        state, ctx.state = ctx.state, None
        in_grads, param_grads = ctx.stage.run_backward(state, grads, ctx.needs_input_grad[2:2 + ctx.n_inputs])
        need_p = ctx.needs_input_grad[2 + ctx.n_inputs:]
        param_grads = tuple(g if n else None for g, n in zip(param_grads, need_p))
        return (None, None) + tuple(in_grads) + param_grads


class _Stage(nn.Module):
    """A module whose forward is one autograd node over libecgmm kernels."""

    def _apply_stage(self, *inputs):
        params = self.cached_params()
        return _StageFn.apply(self, len(inputs), *inputs, *params)

    def stage_params(self):
        return list(self.parameters())

    def _exec_order_params(self):
        return self.stage_params()

    # The parameter list and the arena layout of a stage are walked out of the module tree once and kept
    # (the traversal was ~10 % of the host time of a step).  The cache is dropped by train()/eval() and by
    # .to()/.cuda()/... (every epoch of train.py), so replacing a sub-module takes effect at the next mode switch;
    # call invalidate_caches() to force it.
    def cached_params(self):
        ps = self.__dict__.get("_params_cache")
        if ps is None:
            ps = self.__dict__["_params_cache"] = list(self.stage_params())
        return ps

    def new_arena(self, device):
        lay = self.__dict__.get("_arena_layout")
        if lay is None:
            lay = self.__dict__["_arena_layout"] = ArenaLayout(self._exec_order_params())
        return GradArena(lay, device)

    def invalidate_caches(self):
        self.__dict__.pop("_params_cache", None)
        self.__dict__.pop("_arena_layout", None)

    def train(self, mode=True):
        self.invalidate_caches()
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self.invalidate_caches()
        return super()._apply(fn, *a, **k)

    def precompute(self, *inputs):
        """Enqueue the forward kernels NOW and let the next forward() call only create the autograd node.

        The fusion model uses this to put the image encoder's kernels on the GPU before the ~100 small launches
        of the other encoders are enqueued, while still creating the image node LAST, so that the autograd
        engine (ready nodes run in decreasing creation order) starts backward with the image encoder too."""
        save = torch.is_grad_enabled() and any(t.requires_grad for t in (*inputs, *self.cached_params()))
        with torch.no_grad():
            key = (save, tuple((t.data_ptr(), t._version) for t in inputs))
            self.__dict__["_precomputed"] = (key, self.run_forward(*inputs, save=save))

    def _notify(self, arena, upto):
        """Tell the data-parallel wrapper that arena.flat[:upto] holds final gradients."""
        cb = getattr(self, "_grad_ready_cb", None)
        if cb is not None:
            cb(arena, upto)


# ============================================================================ image encoder
class BasicBlock(nn.Module):
    """torchvision BasicBlock container (resnet.py:59-104)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = BatchNorm2d(planes)
        self.relu = ReLU(inplace=True)
        self.conv2 = Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(Conv2d(inplanes, planes, 1, stride, bias=False), BatchNorm2d(planes))
        self.stride = stride

    def forward(self, x):
        raise lib.EcgmmError("BasicBlock is executed by ResNet18.forward")


def _conv_block_fwd(blk, x, save, one_d=False):
    """conv-BN-ReLU-conv-BN(-SE)(+identity / downsample)-ReLU on channels-last bf16 x."""
    s = blk.stride
    w1f, _ = blk.conv1.shadows()
    a, pa = _conv_bn(ops.conv2d_fwd, blk.bn1, x, w1f, s)
    sa = _bn_stats(blk.bn1, a, blk.conv1.bias, partials=pa)
    m, mask_m = ops.bn_apply(a, sa, relu=True, want_mask=save)
    w2f, _ = blk.conv2.shadows()
    se = getattr(blk, "se", None)
    b, pb = _conv_bn(ops.conv2d_fwd, blk.bn2, m, w2f, 1, want_nsum=se is not None)
    sb = _bn_stats(blk.bn2, b, blk.conv2.bias, want_nsum=se is not None, partials=pb)
    se_rec = None
    gate = None
    if se is not None:
        L = b.shape[1] * b.shape[2]
        pooled, hid, gate = ops.se_fwd(sb.nsum, sb, se.fc[0].weight, se.fc[0].bias, se.fc[2].weight, se.fc[2].bias, L)
        se_rec = (pooled, hid, gate)
    d = sd = None
    if blk.downsample is not None and len(blk.downsample) > 0:
        wdf, _ = blk.downsample[0].shadows()
        d, pd = _conv_bn(ops.conv2d_fwd, blk.downsample[1], x, wdf, s)
        sd = _bn_stats(blk.downsample[1], d, blk.downsample[0].bias, partials=pd)
        idn, _ = ops.bn_apply(d, sd, relu=False)
    else:
        idn = x
    out, mask_out = ops.bn_apply(b, sb, se=gate, res=idn, relu=True, want_mask=save)
    # backward needs the ReLU masks as bits, not the activations (m / out are kept only as conv inputs)
    rec = (x, a, sa, m, b, sb, d, sd, mask_m, mask_out, se_rec) if save else None
    return out, rec


# BatchNorm-backward sums (sum dz, sum dz * xhat) from the epilogue of the data gradient that produces dz
# (ecgmm_conv2d_dgrad_reduce): the separate reduction pass over (x, dz) disappears.  Correct and under test, but OFF by
# default.  Measured on a B200 inside the batch-512 step (profiles/r02e_*, r02h_*, r02p_*):
#   * 64-channel layers (rolling-accumulator kernel, 1152 MMA clocks per tile): the three passes it removes cost 2.8 ms,
#     the epilogue work costs 4.7 ms with one epilogue warp per TMEM quadrant and 3.1 ms with two;
#   * >= 128 channels (a variant of the CTA-pair kernel with a warp-shuffle transposed reduction, not kept): reduction
#     passes -2.5 ms, data gradients +3.5 ms, although a tile holds 9-37 k MMA clocks -- the step runs at the board's
#     power limit (SM clock 1.55-1.65 of 1.965 GHz), where the extra ALU / shuffle work costs clock for everything else.
FUSED_BWD_REDUCE = os.environ.get("ECGMM_FUSED_BWD_REDUCE", "0")


def _fuse_reduce(channels, se):
    if se is not None or FUSED_BWD_REDUCE == "0":
        return False
    return True


def _conv_block_bwd(blk, rec, dout, G, dout_partials=None, next_reduce=None):
    """Returns (gradient w.r.t. the block input, BwdPartials or None).
    dout_partials: the producer of `dout` already reduced it against this block's bn2 (ops.conv2d_dgrad(reduce_for=));
    next_reduce = (bn_x, mask, BNStats) of the BatchNorm that will consume the returned gradient (the previous block's
    bn2): where the library offers it, the last data gradient of this block reduces for it and the partials are
    returned alongside.
    (Weight gradients on their own stream underneath the BatchNorm backward were tried and measured on a B200 --
    profiles/r02a_*: batch 512 131.5 vs 132.4 ms, the weight-gradient class itself 42 vs 34 ms -- and removed.)"""
    x, a, sa, m, b, sb, d, sd, mask_m, mask_out, se_rec = rec
    s = blk.stride
    _, H, W, _ = x.shape
    R, S = blk.conv1.shadows()[0].shape[1:3]
    se = getattr(blk, "se", None)
    gate = None
    se_ctx = None
    if se is not None:
        pooled, hid, gate = se_rec
        L = b.shape[1] * b.shape[2]
        w1, w2 = se.fc[0].weight, se.fc[2].weight

        def se_ctx(p1, p2, split):
            q, dpre2, dpre1 = ops.se_bwd(p1, p2, split, blk.bn2.weight, blk.bn2.bias, w1, w2, hid, gate, L)
            n = gate.shape[0]
            ops.sgemm(dpre2, hid, w2.shape[0], w2.shape[1], n, transA=True, out=G(w2))
            lib.call("ecgmm_colsum", ops._ptr(dpre2), ops._ptr(G(se.fc[2].bias)), n, w2.shape[0], 0, ops._s())
            ops.sgemm(dpre1, pooled, w1.shape[0], w1.shape[1], n, transA=True, out=G(w1))
            lib.call("ecgmm_colsum", ops._ptr(dpre1), ops._ptr(G(se.fc[0].bias)), n, w1.shape[0], 0, ops._s())
            return q

    fuse = _fuse_reduce(a.shape[-1], se)                    # for dm (this block's bn1)
    fuse_in = _fuse_reduce(x.shape[-1], se) and next_reduce is not None  # for dx (the previous block's bn2)
    db_, dz = ops.bn_backward(b, dout, sb, blk.bn2.weight, mask=mask_out, se=gate, se_ctx=se_ctx, want_dz=True,
                              dgamma=G(blk.bn2.weight), dbeta=G(blk.bn2.bias),
                              partials=dout_partials if se is None else None)
    del dout
    ops.conv2d_wgrad(m, db_, G(blk.conv2.weight), R, S, 1)
    _, w2d = blk.conv2.shadows()
    dm, pm = ops.conv2d_dgrad(db_, w2d, (m.shape[1], m.shape[2]), 1, reduce_for=(a, mask_m, sa)) if fuse else \
        (ops.conv2d_dgrad(db_, w2d, (m.shape[1], m.shape[2]), 1), None)
    del db_
    da, _ = ops.bn_backward(a, dm, sa, blk.bn1.weight, mask=mask_m, dgamma=G(blk.bn1.weight), dbeta=G(blk.bn1.bias),
                            partials=pm)
    del dm
    ops.conv2d_wgrad(x, da, G(blk.conv1.weight), R, S, s)
    _, w1d = blk.conv1.shadows()
    if d is not None:
        dsc, dsbn = blk.downsample[0], blk.downsample[1]
        dd, _ = ops.bn_backward(d, dz, sd, dsbn.weight, y=None, dgamma=G(dsbn.weight), dbeta=G(dsbn.bias))
        ops.conv2d_wgrad(x, dd, G(dsc.weight), 1, 1, s)
        dx = ops.conv2d_dgrad(da, w1d, (H, W), s)
        if fuse_in:
            dx, px = ops.conv2d_dgrad(dd, dsc.shadows()[1], (H, W), s, out=dx, accumulate=True, reduce_for=next_reduce)
        else:
            px = None
            ops.conv2d_dgrad(dd, dsc.shadows()[1], (H, W), s, out=dx, accumulate=True)
    elif fuse_in:
        dx, px = ops.conv2d_dgrad(da, w1d, (H, W), s, out=dz, accumulate=True, reduce_for=next_reduce)
    else:
        dx, px = ops.conv2d_dgrad(da, w1d, (H, W), s, out=dz, accumulate=True), None
    return dx, px


class ResNet18(_Stage):
    """Image branch (torchvision resnet18 layout, resnet.py:166-284) as one fused stage.

    forward(image [B,3,H,W] fp32 or bf16, NCHW) -> [B, fc.out_features] fp32."""

    def __init__(self, num_classes=1000):
        super().__init__()
        self.conv1 = StemConv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = BatchNorm2d(64)
        self.relu = ReLU(inplace=True)
        self.maxpool = MaxPool()
        self.layer1 = nn.Sequential(BasicBlock(64, 64), BasicBlock(64, 64))
        self.layer2 = nn.Sequential(BasicBlock(64, 128, 2), BasicBlock(128, 128))
        self.layer3 = nn.Sequential(BasicBlock(128, 256, 2), BasicBlock(256, 256))
        self.layer4 = nn.Sequential(BasicBlock(256, 512, 2), BasicBlock(512, 512))
        self.avgpool = AvgPool()
        self.fc = Linear(512, num_classes)
        for m in self.modules():  # resnet.py:207-212
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def blocks(self):
        return [b for layer in (self.layer1, self.layer2, self.layer3, self.layer4) for b in layer]

    def forward(self, x):
        return self._apply_stage(x)

    def _exec_order_params(self):
        ps = [self.conv1.weight, self.bn1.weight, self.bn1.bias]
        for blk in self.blocks():
            ps += [blk.conv1.weight, blk.bn1.weight, blk.bn1.bias, blk.conv2.weight, blk.bn2.weight, blk.bn2.bias]
            if blk.downsample is not None:
                ps += [blk.downsample[0].weight, blk.downsample[1].weight, blk.downsample[1].bias]
        ps += [self.fc.weight, self.fc.bias]
        return ps

    def run_forward(self, image, save):
        if image.dim() != 4 or image.shape[1] != 3:
            raise lib.EcgmmError(f"image must be [B,3,H,W], got {tuple(image.shape)}")
        if not image.is_cuda:
            raise lib.EcgmmError("image must be a CUDA tensor (no CPU fallback)")
        image = image.detach().contiguous()
        H, W = image.shape[2], image.shape[3]
        _refresh_conv_shadows(self)
        xs = ops.stem_s2d(image)
        ws, _ = self.conv1.shadows()
        if self.bn1.training and FUSED_STATS:
            c1, p1 = ops.stem_conv_fwd(xs, ws, H, W, want_stats=True)
        else:
            c1, p1 = ops.stem_conv_fwd(xs, ws, H, W), None
        st1 = _bn_stats(self.bn1, c1, partials=p1)
        x, arg = ops.bn_relu_maxpool(c1, st1, want_argmax=save)
        recs = []
        for blk in self.blocks():
            x, rec = _conv_block_fwd(blk, x, save)
            recs.append(rec)
        pooled = ops.avgpool_fwd(x)
        feat = ops.linear_fwd(pooled, self.fc.weight, self.fc.bias)
        state = (xs, c1, st1, arg, recs, pooled, tuple(x.shape), (H, W)) if save else None
        # (the pooled stem output lives in recs[0][0]: it is layer1.0's input)
        return feat, state

    def run_backward(self, state, grads, need_in):
        xs, c1, st1, arg, recs, pooled, last_shape, (H, W) = state
        if need_in[0]:
            raise lib.EcgmmError("gradient w.r.t. the input image is not implemented (conv1 has no dgrad path)")
        dfeat = grads[0].contiguous()
        G = self.new_arena(dfeat.device)
        dpooled = ops.linear_bwd(pooled, self.fc.weight, dfeat, dw=G(self.fc.weight), db=G(self.fc.bias))
        dx = ops.avgpool_bwd(dpooled, last_shape)
        blocks = self.blocks()
        stem_out = recs[0][0]  # pooled stem output = input of layer1.0
        px = None
        for i in range(len(blocks) - 1, -1, -1):
            # rec = (x, a, sa, m, b, sb, d, sd, mask_m, mask_out, se_rec): the gradient this block returns is the
            # upstream gradient of the previous block's bn2 (input b, ReLU mask mask_out, statistics sb)
            nr = (recs[i - 1][4], recs[i - 1][9], recs[i - 1][5]) if i > 0 else None
            dx, px = _conv_block_bwd(blocks[i], recs[i], dx, G, dout_partials=px, next_reduce=nr)
            recs[i] = None
            if i in (6, 4, 2):  # a ResNet stage (2 blocks) is complete: its gradients are final
                self._notify(G, G.end_of(blocks[i].conv1.weight))
        dc1, _ = ops.bn_backward(c1, dx, st1, self.bn1.weight, argmax=arg, dgamma=G(self.bn1.weight),
                                 dbeta=G(self.bn1.bias), pooled=stem_out, beta=self.bn1.bias)
        ops.stem_conv_wgrad(xs, dc1, G(self.conv1.weight), H, W)
        self._notify(G, G.total)
        return (None,), [G(p) for p in self.cached_params()]


# ============================================================================ dense leaf modules
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x2 = x.detach().reshape(-1, x.shape[-1]).contiguous()
        y = ops.linear_fwd(x2, w.detach(), None if b is None else b.detach())
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy2 = dy.reshape(-1, w.shape[0]).contiguous()
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        db = torch.empty(w.shape[0], dtype=F32, device=w.device) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        dx = ops.linear_bwd(x2, w.detach(), dy2, need_dx=ctx.needs_input_grad[0], dw=dw, db=db)
        return (dx.view(ctx.xshape) if dx is not None else None), dw, db


class Linear(nn.Linear):
    """nn.Linear container; standalone calls run the libecgmm SGEMM."""

    def forward(self, x):
        _require_cuda_f32(x, "Linear input")
        return _LinearFn.apply(x, self.weight, self.bias)


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, g, b, eps):
        x2 = x.detach().reshape(-1, x.shape[-1]).contiguous()
        y, mean, rstd = ops.layernorm_fwd(x2, g.detach(), b.detach(), eps)
        ctx.save_for_backward(x2, g, mean, rstd)
        ctx.xshape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, g, mean, rstd = ctx.saved_tensors
        dy2 = dy.reshape(x2.shape).contiguous()
        dg = torch.empty_like(g)
        dbt = torch.empty_like(g)
        dx = ops.layernorm_bwd(x2, dy2, g.detach(), mean, rstd, dg, dbt, need_dx=ctx.needs_input_grad[0])
        return (dx.view(ctx.xshape) if dx is not None else None), dg, dbt, None


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        _require_cuda_f32(x, "LayerNorm input")
        return _LayerNormFn.apply(x, self.weight, self.bias, self.eps)


class Dropout(nn.Dropout):
    """Container; `mask_override` (0 or 1/(1-p) per element) injects a mask for parity tests."""
    mask_override = None

    def forward(self, x):
        raise lib.EcgmmError("Dropout is executed by the fused stage that owns it")


def _require_cuda_f32(x, what):
    if not x.is_cuda:
        raise lib.EcgmmError(f"{what} must be a CUDA tensor (no CPU fallback)")
    if x.dtype != F32:
        raise lib.EcgmmError(f"{what} must be float32, got {x.dtype}")


def _drop_fwd(mod, h, training):
    """Dropout on fp32 activations; returns (y, mask or None)."""
    if not training or mod.p == 0.0:
        return h, None
    if mod.mask_override is not None:
        y, mask = ops.dropout_fwd(h, mod.p, 0, mask_in=mod.mask_override.to(h.device, F32).contiguous())
        return y, mask
    return ops.dropout_fwd(h, mod.p, _dropout_seed())


class MLPHead(_Stage, nn.Sequential):
    """Sequential(Linear, ReLU, Dropout, Linear) [optionally with a leading Flatten] as one node.

    Used for `fusion_classifier` (multimodal_paper_modal_balance.py:283-289; indexable, explainers
    read `[0].weight`) and for ResNet1D_SE.classifier (…:112-118)."""

    eval_backward_ok = True  # no batch statistics: explainers differentiate it in eval mode

    def __init__(self, d_in, d_hidden, d_out, p=0.3, flatten=False):
        mods = ([Flatten()] if flatten else []) + [Linear(d_in, d_hidden), ReLU(), Dropout(p), Linear(d_hidden, d_out)]
        nn.Sequential.__init__(self, *mods)
        self._o = 1 if flatten else 0

    @property
    def lin1(self):
        return self[self._o]

    @property
    def drop(self):
        return self[self._o + 2]

    @property
    def lin2(self):
        return self[self._o + 3]

    def forward(self, x):
        _require_cuda_f32(x, "MLP input")
        shape = x.shape
        y = self._apply_stage(x.reshape(-1, shape[-1]))
        return y.view(*shape[:-1], y.shape[-1])

    def stage_params(self):
        return [self.lin1.weight, self.lin1.bias, self.lin2.weight, self.lin2.bias]

    def run_forward(self, x, save):
        x = x.detach().contiguous()
        h = ops.linear_fwd(x, self.lin1.weight, self.lin1.bias, relu=True)
        hd, mask = _drop_fwd(self.drop, h, self.training)
        y = ops.linear_fwd(hd, self.lin2.weight, self.lin2.bias)
        return y, ((x, h, hd, mask) if save else None)

    def mlp_backward(self, state, dy, G, need_dx=True):
        x, h, hd, mask = state
        dhd = ops.linear_bwd(hd, self.lin2.weight, dy, dw=G(self.lin2.weight), db=G(self.lin2.bias))
        dh = ops.mask_bwd(dhd, y=h, mask=mask)
        return ops.linear_bwd(x, self.lin1.weight, dh, need_dx=need_dx, dw=G(self.lin1.weight), db=G(self.lin1.bias))

    def run_backward(self, state, grads, need_in):
        dy = grads[0].contiguous()
        G = self.new_arena(dy.device)
        dx = self.mlp_backward(state, dy, G, need_dx=True)
        return (dx,), [G(p) for p in self.cached_params()]



# ============================================================================ signal encoder
class SEBlock(nn.Module):
    """multimodal_paper_modal_balance.py:49-64 container."""

    def __init__(self, channels, reduction=16):
        super().__init__()
        self.pool = AvgPool()
        self.fc = nn.Sequential(Linear(channels, channels // reduction), ReLU(),
                                Linear(channels // reduction, channels), Sigmoid())

    def forward(self, x):
        raise lib.EcgmmError("SEBlock is executed by ResNet1D_SE.forward")


class BasicBlock1D(nn.Module):
    """multimodal_paper_modal_balance.py:67-93 container."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1):
        super().__init__()
        if kernel_size != 3:
            raise lib.EcgmmError("BasicBlock1D supports kernel_size=3 (the only value the reference uses)")
        self.conv1 = Conv1d(in_channels, out_channels, 3, stride=stride, padding=1)
        self.bn1 = BatchNorm1d(out_channels)
        self.relu = ReLU()
        self.conv2 = Conv1d(out_channels, out_channels, 3, padding=1)
        self.bn2 = BatchNorm1d(out_channels)
        self.se = SEBlock(out_channels)
        self.downsample = None
        if in_channels != out_channels or stride != 1:
            self.downsample = nn.Sequential(Conv1d(in_channels, out_channels, kernel_size=1, stride=stride),
                                            BatchNorm1d(out_channels))
        self.stride = stride

    def forward(self, x):
        raise lib.EcgmmError("BasicBlock1D is executed by ResNet1D_SE.forward")


class ResNet1D_SE(_Stage):
    """Signal branch (multimodal_paper_modal_balance.py:96-125 = signal_model.py:59-88) as one stage.

    forward(x [B,Cin,L] fp32) -> [B,num_classes] fp32.  Conv1d biases are carried by the
    BatchNorm that follows each convolution (training: they only move running_mean, so their
    gradient is identically zero; eval: folded into the BatchNorm shift)."""

    def __init__(self, input_channels=1, num_classes=2, base_filters=64):
        super().__init__()
        if base_filters != 64:
            raise lib.EcgmmError("ResNet1D_SE supports base_filters=64 (the only value the reference uses)")
        self.initial = nn.Sequential(SignalStemConv1d(input_channels, 64, kernel_size=7, stride=2, padding=3),
                                     BatchNorm1d(64), ReLU(), MaxPool())
        self.layer1 = BasicBlock1D(64, 64)
        self.layer2 = BasicBlock1D(64, 128, stride=2)
        self.layer3 = BasicBlock1D(128, 256, stride=2)
        self.global_pool = AvgPool()
        self.classifier = MLPHead(256, 64, num_classes, p=0.3, flatten=True)

    def blocks(self):
        return [self.layer1, self.layer2, self.layer3]

    def forward(self, x):
        return self._apply_stage(x)

    def _exec_order_params(self):
        ps = [self.initial[0].weight, self.initial[0].bias, self.initial[1].weight, self.initial[1].bias]
        for blk in self.blocks():
            ps += [blk.conv1.weight, blk.conv1.bias, blk.bn1.weight, blk.bn1.bias, blk.conv2.weight, blk.conv2.bias,
                   blk.bn2.weight, blk.bn2.bias, blk.se.fc[0].weight, blk.se.fc[0].bias, blk.se.fc[2].weight,
                   blk.se.fc[2].bias]
            if blk.downsample is not None:
                ps += [blk.downsample[0].weight, blk.downsample[0].bias, blk.downsample[1].weight,
                       blk.downsample[1].bias]
        ps += self.classifier.stage_params()
        return ps

    def run_forward(self, sig, save):
        if sig.dim() != 3:
            raise lib.EcgmmError(f"signal must be [B,C,L], got {tuple(sig.shape)}")
        if not sig.is_cuda:
            raise lib.EcgmmError("signal must be a CUDA tensor (no CPU fallback)")
        sig = sig.detach().to(F32).contiguous()
        _refresh_conv_shadows(self)
        stem, bn0 = self.initial[0], self.initial[1]
        # tensor-core stem (4 samples of all leads per 64-channel pixel) where the length allows it, else the direct kernels
        xs4 = ops.signal_s4d(sig) if SIGNAL_STEM_TC else None
        if xs4 is not None:
            c0 = ops.conv2d_fwd(xs4, stem.shadows()[0], 1).view(sig.shape[0], 1, -1, 64)
            sig = xs4  # what backward needs
        else:
            c0 = ops.signal_stem_fwd(sig, stem.weight.detach())
        st0 = _bn_stats(bn0, c0, stem.bias)
        x, arg = ops.bn_relu_maxpool(c0, st0, want_argmax=save)
        recs = []
        for blk in self.blocks():
            x, rec = _conv_block_fwd(blk, x, save, one_d=True)
            recs.append(rec)
        pooled = ops.avgpool_fwd(x)
        y, mlp_state = self.classifier.run_forward(pooled, save)
        state = (sig, c0, st0, arg, recs, tuple(x.shape), mlp_state) if save else None
        return y, state

    def run_backward(self, state, grads, need_in):
        sig, c0, st0, arg, recs, last_shape, mlp_state = state
        if need_in[0]:
            raise lib.EcgmmError("gradient w.r.t. the input signal is not implemented")
        dy = grads[0].contiguous()
        G = self.new_arena(dy.device)
        dpooled = self.classifier.mlp_backward(mlp_state, dy, G)
        dx = ops.avgpool_bwd(dpooled, last_shape)
        blocks = self.blocks()
        stem_out = recs[0][0]
        for i in range(len(blocks) - 1, -1, -1):
            dx, _ = _conv_block_bwd(blocks[i], recs[i], dx, G)
            recs[i] = None
        bn0 = self.initial[1]
        dc0, _ = ops.bn_backward(c0, dx, st0, bn0.weight, argmax=arg, dgamma=G(bn0.weight), dbeta=G(bn0.bias),
                                 pooled=stem_out, beta=bn0.bias)
        if sig.dtype == BF16:  # the regrouped signal of the tensor-core stem
            ops.signal_stem_wgrad_s4d(sig, dc0, G(self.initial[0].weight))
        else:
            ops.signal_stem_wgrad(sig, dc0, G(self.initial[0].weight))
        self._notify(G, G.total)
        return (None,), [G(p) for p in self.cached_params()]


# ============================================================================ clinical encoder
class ClinicalMLP(_Stage, nn.Sequential):
    """Linear(F,64) -> BatchNorm1d(64) -> ReLU -> Dropout(0.3) -> Linear(64,D)
    (multimodal_paper_modal_balance.py:256-262), one node."""

    def __init__(self, n_features, d_out, hidden=64, p=0.3):
        nn.Sequential.__init__(self, Linear(n_features, hidden), BatchNorm1d(hidden), ReLU(), Dropout(p),
                               Linear(hidden, d_out))

    def forward(self, x):
        _require_cuda_f32(x, "clinical input")
        return self._apply_stage(x)

    def stage_params(self):
        return [self[0].weight, self[0].bias, self[1].weight, self[1].bias, self[4].weight, self[4].bias]

    def run_forward(self, x, save):
        x = x.detach().contiguous()
        bn = self[1]
        B, C = x.shape[0], bn.num_features
        z = ops.linear_fwd(x, self[0].weight, self[0].bias)
        y = torch.empty_like(z)
        st = torch.empty(2 * C, dtype=F32, device=x.device)
        mean, invstd = st[:C], st[C:]
        lib.call("ecgmm_bn_rows_fwd", ops._ptr(z), ops._ptr(bn.weight), ops._ptr(bn.bias), ops._ptr(bn.running_mean),
                 ops._ptr(bn.running_var), ops._ptr(bn.num_batches_tracked), ops._ptr(y), ops._ptr(mean),
                 ops._ptr(invstd), B, C, float(bn.eps), float(bn.momentum), int(bn.training), 1, ops._s())
        yd, mask = _drop_fwd(self[3], y, self.training)
        out = ops.linear_fwd(yd, self[4].weight, self[4].bias)
        return out, ((x, z, y, yd, mask, mean, invstd) if save else None)

    def run_backward(self, state, grads, need_in):
        x, z, y, yd, mask, mean, invstd = state
        dout = grads[0].contiguous()
        G = self.new_arena(dout.device)
        bn = self[1]
        B, C = z.shape
        dyd = ops.linear_bwd(yd, self[4].weight, dout, dw=G(self[4].weight), db=G(self[4].bias))
        dy = ops.mask_bwd(dyd, y=None, mask=mask) if mask is not None else dyd
        dz = torch.empty_like(z)
        lib.call("ecgmm_bn_rows_bwd", ops._ptr(z), ops._ptr(dy), ops._ptr(y), ops._ptr(bn.weight), ops._ptr(mean),
                 ops._ptr(invstd), ops._ptr(dz), ops._ptr(G(bn.weight)), ops._ptr(G(bn.bias)), B, C, 1, ops._s())
        dx = ops.linear_bwd(x, self[0].weight, dz, need_dx=need_in[0], dw=G(self[0].weight), db=G(self[0].bias))
        self._notify(G, G.total)
        return (dx,), [G(p) for p in self.cached_params()]


# ============================================================================ fusion head
class _GateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f0, f1, f2, weights, norm_w, norm_b, eps):
        f0, f1, f2 = (t.detach().contiguous() for t in (f0, f1, f2))
        B, D0, D1, D2 = f0.shape[0], f0.shape[1], f1.shape[1], f2.shape[1]
        fused_pre = torch.empty((B, D0 + D1 + D2), dtype=F32, device=f0.device)
        soft_w = torch.empty(3, dtype=F32, device=f0.device)
        lib.call("ecgmm_fusion_gate_fwd", ops._ptr(f0), ops._ptr(f1), ops._ptr(f2), ops._ptr(weights.detach()),
                 ops._ptr(fused_pre), ops._ptr(soft_w), B, D0, D1, D2, ops._s())
        fused, mean, rstd = ops.layernorm_fwd(fused_pre, norm_w.detach(), norm_b.detach(), eps)
        ctx.save_for_backward(f0, f1, f2, weights, norm_w, fused_pre, mean, rstd)
        ctx.mark_non_differentiable(soft_w)
        return fused, soft_w

    @staticmethod
    def backward(ctx, dfused, _dsw):
        f0, f1, f2, weights, norm_w, fused_pre, mean, rstd = ctx.saved_tensors
        dnw, dnb = torch.empty_like(norm_w), torch.empty_like(norm_w)
        dpre = ops.layernorm_bwd(fused_pre, dfused.contiguous(), norm_w.detach(), mean, rstd, dnw, dnb)
        df0, df1, df2 = torch.empty_like(f0), torch.empty_like(f1), torch.empty_like(f2)
        dwt = torch.empty(3, dtype=F32, device=f0.device)
        lib.call("ecgmm_fusion_gate_bwd", ops._ptr(dpre), ops._ptr(f0), ops._ptr(f1), ops._ptr(f2),
                 ops._ptr(weights.detach()), ops._ptr(df0), ops._ptr(df1), ops._ptr(df2), ops._ptr(dwt), f0.shape[0],
                 f0.shape[1], f1.shape[1], f2.shape[1], 0, ops._s())
        return df0, df1, df2, dwt, dnw, dnb, None


class AttentionFusion(nn.Module):
    """multimodal_paper_modal_balance.py:31-46: softmax over 3 scalars, scale, concat, LayerNorm.
    The soft weights are returned for logging only (non-differentiable, as used by train.py:126-133)."""

    def __init__(self, dims):
        super().__init__()
        self.weights = nn.Parameter(torch.ones(3))
        self.norm = LayerNorm(sum(dims))

    def forward(self, img_feat, signal_feat, clinical_feat):
        for t in (img_feat, signal_feat, clinical_feat):
            _require_cuda_f32(t, "AttentionFusion input")
        return _GateFn.apply(img_feat, signal_feat, clinical_feat, self.weights, self.norm.weight, self.norm.bias,
                             self.norm.eps)


class FusionHead(_Stage):
    """Everything after the three encoders in ECGMultimodalModel.forward
    (multimodal_paper_modal_balance.py:326,329,332,336-352) as ONE autograd node: the three
    LayerNorms, the three branch classifiers, AttentionFusion, fusion_classifier and var_loss.
    It owns no parameters: it borrows the sub-modules of the model so the state_dict layout of
    the reference is untouched."""

    non_differentiable_outputs = (5,)
    eval_backward_ok = True

    def __init__(self, model):
        super().__init__()
        object.__setattr__(self, "_m", model)  # not registered: avoids a module cycle / duplicate keys

    def stage_params(self):
        m = self._m
        fc = m.fusion_classifier
        return [m.image_norm.weight, m.image_norm.bias, m.signal_norm.weight, m.signal_norm.bias,
                m.clinical_norm.weight, m.clinical_norm.bias,
                m.image_classifier.weight, m.image_classifier.bias, m.signal_classifier.weight,
                m.signal_classifier.bias, m.clinical_classifier.weight, m.clinical_classifier.bias,
                m.attention_fusion.weights, m.attention_fusion.norm.weight, m.attention_fusion.norm.bias,
                fc.lin1.weight, fc.lin1.bias, fc.lin2.weight, fc.lin2.bias]

    @property
    def training(self):
        return self._m.training

    @training.setter
    def training(self, v):
        pass

    def forward(self, e0, e1, e2):
        return self._apply_stage(e0, e1, e2)

    def run_forward(self, e0, e1, e2, save):
        m = self._m
        encs = [t.detach().to(F32).contiguous() for t in (e0, e1, e2)]
        norms = (m.image_norm, m.signal_norm, m.clinical_norm)
        heads = (m.image_classifier, m.signal_classifier, m.clinical_classifier)
        feats, lnst, logits = [], [], []
        for e, n, h in zip(encs, norms, heads):
            f, mean, rstd = ops.layernorm_fwd(e, n.weight, n.bias, n.eps)
            feats.append(f)
            lnst.append((mean, rstd))
            logits.append(ops.linear_fwd(f, h.weight, h.bias))
        B = encs[0].shape[0]
        D = [f.shape[1] for f in feats]
        dev = encs[0].device
        af = m.attention_fusion
        fused_pre = torch.empty((B, sum(D)), dtype=F32, device=dev)
        soft_w = torch.empty(3, dtype=F32, device=dev)
        lib.call("ecgmm_fusion_gate_fwd", ops._ptr(feats[0]), ops._ptr(feats[1]), ops._ptr(feats[2]),
                 ops._ptr(af.weights), ops._ptr(fused_pre), ops._ptr(soft_w), B, D[0], D[1], D[2], ops._s())
        fused, fmean, frstd = ops.layernorm_fwd(fused_pre, af.norm.weight, af.norm.bias, af.norm.eps)
        flogits, mlp_state = m.fusion_classifier.run_forward(fused, save)
        small = torch.empty(4 + 3 * B, dtype=F32, device=dev)
        var_loss, coef, row_mean = small[0:1], small[1:4], small[4:]
        lib.call("ecgmm_var_loss_fwd", ops._ptr(feats[0]), ops._ptr(feats[1]), ops._ptr(feats[2]), ops._ptr(var_loss),
                 ops._ptr(row_mean), ops._ptr(coef), B, D[0], D[1], D[2], ops._s())
        state = (encs, feats, lnst, fused_pre, fmean, frstd, mlp_state, coef, row_mean) if save else None
        return (logits[0], logits[1], logits[2], flogits, var_loss.view(()), soft_w), state

    def run_backward(self, state, grads, need_in):
        m = self._m
        encs, feats, lnst, fused_pre, fmean, frstd, mlp_state, coef, row_mean = state
        dl = grads[0:3]
        dlf, dvar = grads[3], grads[4]
        params = self.cached_params()
        G = self.new_arena(encs[0].device)
        B = encs[0].shape[0]
        D = [f.shape[1] for f in feats]
        heads = (m.image_classifier, m.signal_classifier, m.clinical_classifier)
        norms = (m.image_norm, m.signal_norm, m.clinical_norm)
        unused = set()
        dfeat = []
        for i in range(3):
            if dl[i] is not None:
                dfeat.append(ops.linear_bwd(feats[i], heads[i].weight, dl[i].contiguous(), dw=G(heads[i].weight),
                                            db=G(heads[i].bias)))
            else:
                dfeat.append(torch.zeros_like(feats[i]))
                unused.update((id(heads[i].weight), id(heads[i].bias)))
        af = m.attention_fusion
        fc = m.fusion_classifier
        if dlf is not None:
            dfused = fc.mlp_backward(mlp_state, dlf.contiguous(), G)
            dpre = ops.layernorm_bwd(fused_pre, dfused, af.norm.weight, fmean, frstd, G(af.norm.weight),
                                     G(af.norm.bias))
            lib.call("ecgmm_fusion_gate_bwd", ops._ptr(dpre), ops._ptr(feats[0]), ops._ptr(feats[1]),
                     ops._ptr(feats[2]), ops._ptr(af.weights), ops._ptr(dfeat[0]), ops._ptr(dfeat[1]),
                     ops._ptr(dfeat[2]), ops._ptr(G(af.weights)), B, D[0], D[1], D[2], 1, ops._s())
        else:
            unused.update(id(p) for p in (af.weights, af.norm.weight, af.norm.bias, *fc.stage_params()))
        if dvar is not None:
            gv = dvar.detach().to(F32).reshape(1).contiguous()
            for i in range(3):
                lib.call("ecgmm_var_loss_bwd", ops._ptr(feats[i]), ops._ptr(row_mean[i * B:(i + 1) * B]),
                         ops._ptr(coef[i:i + 1]), ops._ptr(gv), ops._ptr(dfeat[i]), B, D[i], 1, ops._s())
        denc = []
        for i in range(3):
            mean, rstd = lnst[i]
            denc.append(ops.layernorm_bwd(encs[i], dfeat[i], norms[i].weight, mean, rstd, G(norms[i].weight),
                                          G(norms[i].bias), need_dx=need_in[i]))
        self._notify(G, G.total)
        return tuple(denc), [None if id(p) in unused else G(p) for p in params]


# ============================================================================ the fusion model
class DefaultConfig:
    """Subset of the reference Config (config.py:6-46) the model reads."""
    num_classes = 2
    device = "cuda"


class ECGMultimodalModel(nn.Module):
    """Drop-in for multimodal_paper_modal_balance.ECGMultimodalModel (…:197-354).

    ECGMultimodalModel(config) reads config.num_classes / config.device like the reference.
    forward(image [B,3,H,W], ecg_signal [B,L] or [B,C,L], clinical [B,F]) ->
        (img_logits, signal_logits, clinical_logits, fusion_logits, var_loss, soft_weights).
    Extra keyword options (reference-compatible defaults): dims=(256,256,256) (G2) or
    (512,128,32) (G3 layout), clinical_features=24, signal_channels=1, fusion_only=False
    (True returns only fusion_logits, the single-tensor API train_kfold.py:59-64 expects).
    Unlike the reference constructor it does not read checkpoint files from ./checkpoints;
    use load_pretrained_image_encoder / load_pretrained_signal_encoder."""

    def __init__(self, config=None, dims=(256, 256, 256), clinical_features=24, signal_channels=1,
                 fusion_only=False):
        super().__init__()
        config = config if config is not None else DefaultConfig
        num_classes = getattr(config, "num_classes", 2)
        self.image_dim, self.signal_dim, self.clinical_dim = dims
        self.modal_dim = self.image_dim
        self.fusion_only = fusion_only
        self._clinical_features = clinical_features
        self.image_encoder = ResNet18()
        self.image_encoder.fc = Linear(512, self.image_dim)
        self.image_norm = LayerNorm(self.image_dim)
        self.signal_encoder = ResNet1D_SE(input_channels=signal_channels, num_classes=self.signal_dim)
        self.signal_norm = LayerNorm(self.signal_dim)
        self.clinical_encoder = ClinicalMLP(clinical_features, self.clinical_dim)
        self.clinical_norm = LayerNorm(self.clinical_dim)
        self.image_classifier = Linear(self.image_dim, num_classes)
        self.signal_classifier = Linear(self.signal_dim, num_classes)
        self.clinical_classifier = Linear(self.clinical_dim, num_classes)
        self.attention_fusion = AttentionFusion(dims=list(dims))
        self.fusion_classifier = MLPHead(sum(dims), 128, num_classes, p=0.3)
        object.__setattr__(self, "_head", FusionHead(self))
        object.__setattr__(self, "_streams", {})
        import os as _os

        self.overlap_branches = _os.environ.get("ECGMM_SIDE_STREAM", "1") != "0"
        dev = getattr(config, "device", None)
        if dev is not None and str(dev) != "cpu" and torch.cuda.is_available():
            self.to(dev)

    def get_clinical_feature_dim(self):
        return self._clinical_features

    # the fusion head borrows this module's children without being registered as one: forward the cache resets
    def train(self, mode=True):
        self._head.invalidate_caches()
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self._head.invalidate_caches()
        return super()._apply(fn, *a, **k)

    def stages(self):
        return [self.image_encoder, self.signal_encoder, self.clinical_encoder, self._head]

    def forward(self, image, ecg_signal, clinical):
        if ecg_signal.dim() == 2:
            ecg_signal = ecg_signal.unsqueeze(1)  # multimodal_paper_modal_balance.py:328
        if self.overlap_branches and image.is_cuda:
            # The signal and clinical encoders are ~300 small latency-bound launches that do not depend on
            # the image encoder: they run on a side stream (forward here, and - because autograd replays a
            # node's backward on its forward stream - backward too) underneath the big conv / BN kernels.
            main = torch.cuda.current_stream()
            side = self._side_stream()
            side.wait_stream(main)  # inputs (and the previous step's use of recycled buffers) are ready
            self.image_encoder.precompute(image)  # GPU starts on the convolutions while the host enqueues the rest
            with torch.cuda.stream(side):
                e_sig = self.signal_encoder(ecg_signal)
                e_clin = self.clinical_encoder(clinical)
            e_img = self.image_encoder(image)
            main.wait_stream(side)
            for t in (ecg_signal, clinical, e_sig, e_clin):
                t.record_stream(main if t is e_sig or t is e_clin else side)
        else:
            e_img = self.image_encoder(image)
            e_sig = self.signal_encoder(ecg_signal)
            e_clin = self.clinical_encoder(clinical)
        out = self._head(e_img, e_sig, e_clin)
        return out[3] if self.fusion_only else out

    def _side_stream(self):
        dev = torch.cuda.current_device()
        st = self._streams.get(dev)
        if st is None:
            st = self._streams[dev] = torch.cuda.Stream()
        return st

    # ---- checkpoint helpers (multimodal_paper_modal_balance.py:291-322,356-383)
    def load_pretrained_signal_encoder(self, path, load_fc=False):
        sd = torch.load(path, map_location="cpu")
        if not load_fc:
            sd = {k: v for k, v in sd.items() if not k.startswith("classifier.4.")}
        return self.signal_encoder.load_state_dict(sd, strict=False)

    def load_pretrained_image_encoder(self, path, load_fc=False):
        sd = torch.load(path, map_location="cpu")
        if not load_fc:
            sd = {k: v for k, v in sd.items() if not k.startswith("fc.")}
        return self.image_encoder.load_state_dict(sd, strict=False)


MultimodalModel = ECGMultimodalModel  # name used by BASELINE.json's north_star


class FusionClassifierWrapper(nn.Module):
    """fusion_classifier.py:5-11 -- exposes the fusion head alone for SHAP / LIME drivers."""

    def __init__(self, fusion_classifier):
        super().__init__()
        self.fusion_classifier = fusion_classifier

    def forward(self, x):
        return self.fusion_classifier(x)

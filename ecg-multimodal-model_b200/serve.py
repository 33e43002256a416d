"""Eval / serving path of the image branch.

The reference has no server code; its mobile client posts one ECG image to an endpoint
(Groove/components/SubmitButton.tsx:44-45) and shows a class with a heat map (gpt/*.png).  What that endpoint has to
run is the image-only chain of the fusion model in eval mode,

    image_encoder -> image_norm -> image_classifier          (multimodal_paper_modal_balance.py:325-327,337)

followed by softmax / argmax, and Grad-CAM on the last ResNet stage:

    cam[n, y, x] = relu( sum_k alpha[n, k] * A[n, y, x, k] ),   alpha[n, k] = mean_{y,x} d logit_c / d A[n, y, x, k]

Everything runs through libecgmm: BatchNorm uses its frozen statistics as a per-channel scale / shift (no statistics
pass), dropout does not exist on this path, raw uint8 pixels are normalised inside the first kernel, and for a fixed
batch shape the whole request is ONE CUDA-graph launch (`graph=True`).  The Grad-CAM gradient needs no backward pass
through the convolutions: behind layer4 there are only the average pool (uniform gradient 1/(h w)), fc, LayerNorm and
the classifier row of the requested class, i.e. one row gather, one LayerNorm backward and one [N,256]x[256,512] SGEMM.

    ep = ecgmm.serve.ImageEndpoint(model, example_image=batch_u8)      # model.eval() first
    probs, classes = ep(batch_u8)
    probs, classes, cam = ep.gradcam(batch_u8)                         # cam [N, h, w] fp32 on the layer4 grid
"""
from __future__ import annotations

import torch

import os

from . import graph as graph_mod
from . import lib, ops

F32 = torch.float32
# BatchNorm (+ residual, + ReLU) applied by the convolution epilogue (ecgmm_conv2d_fwd_bn): one kernel and one write per
# conv instead of conv -> scale/shift pass (measured on a B200, batch 64 at 250x2500: 7.02 -> 6.64 ms per request).
# ECGMM_SERVE_FUSED=0 selects the two-kernel path.
FUSED_EPILOGUE = os.environ.get("ECGMM_SERVE_FUSED", "1") != "0"


def fold_batchnorm(model):
    """Frozen BatchNorm statistics of the image branch as per-channel (scale, shift) pairs, computed once
    (ecgmm_bn_eval_coeffs: scale = gamma / sqrt(var + eps), shift = beta - mean * scale).  Returns {id(bn): BNStats}."""
    enc = model.image_encoder
    bns = [enc.bn1]
    for blk in enc.blocks():
        bns += [blk.bn1, blk.bn2] + ([blk.downsample[1]] if blk.downsample is not None else [])
    return {id(bn): ops.bn_eval_coeffs(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps) for bn in bns}


def _block_eval(blk, x, co):
    """torchvision BasicBlock (resnet.py:59-104) with folded statistics: conv -> scale/shift+ReLU -> conv ->
    scale/shift (+ identity or its 1x1 projection) -> ReLU."""
    w1f, _ = blk.conv1.shadows()
    if FUSED_EPILOGUE:
        m = ops.conv2d_fwd_bn(x, w1f, co[id(blk.bn1)], blk.stride, relu=True)
        idn = x
        if blk.downsample is not None:
            wdf, _ = blk.downsample[0].shadows()
            idn = ops.conv2d_fwd_bn(x, wdf, co[id(blk.downsample[1])], blk.stride, relu=False)
        w2f, _ = blk.conv2.shadows()
        return ops.conv2d_fwd_bn(m, w2f, co[id(blk.bn2)], 1, res=idn, relu=True)
    m, _ = ops.bn_apply(ops.conv2d_fwd(x, w1f, blk.stride), co[id(blk.bn1)], relu=True)
    w2f, _ = blk.conv2.shadows()
    b = ops.conv2d_fwd(m, w2f, 1)
    idn = x
    if blk.downsample is not None:
        wdf, _ = blk.downsample[0].shadows()
        idn, _ = ops.bn_apply(ops.conv2d_fwd(x, wdf, blk.stride), co[id(blk.downsample[1])], relu=False)
    out, _ = ops.bn_apply(b, co[id(blk.bn2)], res=idn, relu=True)
    return out


def image_features(model, image, coeffs=None):
    """Eval-mode image branch.  Returns (act [N,h,w,512] bf16 = layer4 output, pooled [N,512], feat [N,D_img],
    normed [N,D_img], ln_mean, ln_rstd, logits [N,C]).  coeffs: fold_batchnorm(model) (computed here if absent)."""
    enc = model.image_encoder
    if enc.training or model.image_norm.training:
        raise lib.EcgmmError("ImageEndpoint serves eval-mode statistics: call model.eval() first")
    if image.dim() != 4 or image.shape[1] != 3:
        raise lib.EcgmmError(f"image must be [B,3,H,W], got {tuple(image.shape)}")
    if not image.is_cuda:
        raise lib.EcgmmError("image must be a CUDA tensor (no CPU fallback)")
    co = coeffs if coeffs is not None else fold_batchnorm(model)
    image = image.detach().contiguous()
    H, W = image.shape[2], image.shape[3]
    xs = ops.stem_s2d(image)
    ws, _ = enc.conv1.shadows()
    c1 = ops.stem_conv_fwd(xs, ws, H, W)
    x, _ = ops.bn_relu_maxpool(c1, co[id(enc.bn1)], want_argmax=False)
    for blk in enc.blocks():
        x = _block_eval(blk, x, co)
    pooled = ops.avgpool_fwd(x)
    feat = ops.linear_fwd(pooled, enc.fc.weight.detach(), enc.fc.bias.detach())
    ln = model.image_norm
    normed, mean, rstd = ops.layernorm_fwd(feat, ln.weight.detach(), ln.bias.detach(), ln.eps)
    cl = model.image_classifier
    logits = ops.linear_fwd(normed, cl.weight.detach(), cl.bias.detach())
    return x, pooled, feat, normed, mean, rstd, logits


def softmax_rows(logits, want_argmax=True):
    """logits [rows, C] fp32 -> (probs [rows, C], argmax [rows] int32 or None)."""
    ops._chk(logits, F32, "logits")
    rows, C = logits.shape
    probs = torch.empty_like(logits)
    am = torch.empty((rows,), dtype=torch.int32, device=logits.device) if want_argmax else None
    lib.call("ecgmm_softmax_rows", ops._ptr(logits), ops._ptr(probs), ops._ptr(am), rows, C, ops._s())
    return probs, am


def gradcam_from_features(model, act, feat, mean, rstd, classes):
    """cam [N, h, w] fp32 for the class of every sample (classes int32 [N] on the device)."""
    N, h, w, C = act.shape
    cl, ln, fc = model.image_classifier, model.image_norm, model.image_encoder.fc
    wc = cl.weight.detach().contiguous()
    dnorm = torch.empty((N, wc.shape[1]), dtype=F32, device=act.device)       # d logit_c / d normed = classifier row c
    lib.call("ecgmm_gather_rows", ops._ptr(wc), ops._ptr(classes), ops._ptr(dnorm), N, wc.shape[1], wc.shape[0],
             ops._s())
    dfeat = ops.layernorm_bwd(feat, dnorm, ln.weight.detach(), mean, rstd)
    dpooled = ops.sgemm(dfeat, fc.weight.detach().contiguous(), N, fc.weight.shape[1], fc.weight.shape[0])
    cam = torch.empty((N, h, w), dtype=F32, device=act.device)
    lib.call("ecgmm_gradcam", ops._ptr(act), ops._ptr(dpooled), ops._ptr(cam), N, h * w, C, 1.0 / float(h * w),
             ops._s())
    return cam


class ImageEndpoint:
    """Callable endpoint over a model in eval mode.

    example_image fixes the request shape and dtype (uint8 raw pixels, bf16 or fp32 normalised); with graph=True the
    request (and, separately, the request + Grad-CAM) is captured once as a CUDA graph and replayed: one launch per
    call.  The graphs are re-captured when a parameter or BatchNorm buffer of the image branch has changed since."""

    def __init__(self, model, example_image=None, graph=True, class_index=None):
        self.model = model
        self.class_index = None if class_index is None else int(class_index)
        self.use_graph = bool(graph) and example_image is not None
        self._graphs = {}
        self._static = None
        self._folded = None  # (versions, {id(bn): BNStats})
        if example_image is not None:
            if not example_image.is_cuda:
                raise lib.EcgmmError("example_image must be a CUDA tensor (no CPU fallback)")
            self._static = example_image.detach().clone().contiguous()

    # ---- the two request bodies (kernel sequences)
    def _classify(self, image):
        _, _, _, _, _, _, logits = image_features(self.model, image, self._coeffs())
        return softmax_rows(logits)

    def _classify_cam(self, image):
        act, _, feat, _, mean, rstd, logits = image_features(self.model, image, self._coeffs())
        probs, am = softmax_rows(logits)
        classes = am
        if self.class_index is not None:
            if not 0 <= self.class_index < logits.shape[1]:
                raise lib.EcgmmError(f"class index {self.class_index} out of range for {logits.shape[1]} classes")
            classes = torch.full_like(am, self.class_index)
        return probs, classes, gradcam_from_features(self.model, act, feat, mean, rstd, classes)

    # ---- folded statistics / graph plumbing
    def _versions(self):
        return tuple((t.data_ptr(), t._version) for t in self._tracked())

    def _coeffs(self):
        """BatchNorm scale / shift pairs, recomputed only when a tensor of the image branch has changed."""
        if self.model.image_encoder.training or self.model.image_norm.training:
            raise lib.EcgmmError("ImageEndpoint serves eval-mode statistics: call model.eval() first")
        vers = self._versions()
        if self._folded is None or self._folded[0] != vers:
            self._folded = (vers, fold_batchnorm(self.model))
        return self._folded[1]

    def _tracked(self):
        m = self.model
        mods = (m.image_encoder, m.image_norm, m.image_classifier)
        return [t for mod in mods for t in list(mod.parameters()) + list(mod.buffers())]

    def _graphed(self, key, body):
        vers = self._versions()
        ent = self._graphs.get(key)
        if ent is None or ent[2] != vers:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side), torch.no_grad():  # warm-up: allocator, weight shadows, kernel attributes
                body(self._static)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with graph_mod._quiet_gc(), torch.no_grad(), torch.cuda.graph(g):
                outs = body(self._static)
            ent = self._graphs[key] = (g, outs, vers)
        return ent

    def _run(self, key, body, image):
        if not self.use_graph:
            with torch.no_grad():
                return body(image)
        if image.shape != self._static.shape or image.dtype != self._static.dtype:
            raise lib.EcgmmError(f"request {tuple(image.shape)} {image.dtype} does not match the captured "
                                 f"{tuple(self._static.shape)} {self._static.dtype}")
        g, outs, _ = self._graphed(key, body)
        if image is not self._static:
            self._static.copy_(image, non_blocking=True)
        g.replay()
        return outs

    @property
    def input(self):
        """The graph's own input buffer: fill it directly (e.g. from a copy stream) and call ep(ep.input)."""
        return self._static

    def __call__(self, image):
        """image [N,3,H,W] -> (probs [N,C] fp32, classes [N] int32).  With graph=True the returned tensors are the
        graph's output buffers: they are overwritten by the next call."""
        return self._run("classify", self._classify, image)

    def gradcam(self, image):
        """image -> (probs [N,C], classes [N] int32, cam [N,h,w] fp32) for class_index (None: each sample's argmax)."""
        return self._run("cam", self._classify_cam, image)

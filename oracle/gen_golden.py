"""Pins the oracle against the REAL reference and writes tests/golden/*.

Run in the build container only (needs /root/reference, which never travels to the GPU box):

    python oracle/gen_golden.py

Steps
 1. import /root/reference/multimodal_paper_modal_balance.py (its constructor reads two
    checkpoint files relative to CWD, so we chdir into a temp dir holding stand-ins);
 2. copy one procedural state_dict (tests/golden_util.py) into the reference model and into
    oracle.model.ECGMultimodalModel and assert BIT-IDENTICAL outputs, loss and gradients in
    eval mode, in train mode with dropout (same torch RNG state), and after one Adam step;
 3. check the known answer of the reference's only real checkpoint (best_ptbxl.pth);
 4. store small golden tensors + checksums under tests/golden/.
"""
import os
import shutil
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"

from golden_util import GOLDEN_DIR, make_inputs, make_oracle, set_dropout, state_checksums  # noqa: E402
from oracle import model as oracle_model  # noqa: E402


def import_reference():
    tmp = tempfile.mkdtemp(prefix="ecgmm_ref_")
    import torchvision

    os.makedirs(os.path.join(tmp, "checkpoints/0716_111810"))
    os.makedirs(os.path.join(tmp, "checkpoints/0716_172631"))
    torch.save(torchvision.models.resnet18().state_dict(), os.path.join(tmp, "checkpoints/0716_111810/last.pth"))
    shutil.copy(os.path.join(REF, "best_ptbxl.pth"), os.path.join(tmp, "checkpoints/0716_172631/best.pth"))
    cwd = os.getcwd()
    os.chdir(tmp)
    sys.path.insert(0, REF)
    import multimodal_paper_modal_balance as ref  # noqa
    from config import Config  # noqa

    Config.device = "cpu"
    model = ref.ECGMultimodalModel(Config)
    os.chdir(cwd)
    return ref, model


def assert_same(a, b, what):
    if isinstance(a, (tuple, list)):
        for i, (x, y) in enumerate(zip(a, b)):
            assert_same(x, y, f"{what}[{i}]")
        return
    assert torch.equal(a, b), f"oracle != reference for {what}: max diff {(a - b).abs().max().item()}"


def run_pair(ref_m, ora_m, inputs, mode, seed):
    image, ecg, clin, labels = inputs
    outs = []
    for m in (ref_m, ora_m):
        m.train(mode == "train")
        m.zero_grad(set_to_none=True)
        torch.manual_seed(seed)  # same dropout masks
        o = m(image, ecg, clin)
        loss = oracle_model.fusion_loss(o, labels)
        loss.backward()
        outs.append((o, loss.detach(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    assert_same(outs[0][0], outs[1][0], f"{mode} outputs")
    assert_same(outs[0][1], outs[1][1], f"{mode} loss")
    assert outs[0][2].keys() == outs[1][2].keys()
    for k in outs[0][2]:
        assert_same(outs[0][2][k], outs[1][2][k], f"{mode} grad {k}")
    return outs[1]


SMALL_GRAD_KEYS = [
    "attention_fusion.weights", "attention_fusion.norm.weight", "fusion_classifier.0.bias",
    "fusion_classifier.3.weight", "fusion_classifier.3.bias", "image_norm.weight", "signal_norm.bias",
    "clinical_norm.weight", "image_encoder.fc.bias", "image_encoder.bn1.weight", "image_encoder.bn1.bias",
    "image_encoder.layer1.0.bn2.weight", "image_encoder.layer4.1.bn2.bias", "image_encoder.layer2.0.downsample.1.weight",
    "signal_encoder.initial.1.weight", "signal_encoder.layer2.bn2.bias", "signal_encoder.layer3.se.fc.0.bias",
    "signal_encoder.classifier.1.bias", "clinical_encoder.1.weight", "clinical_encoder.4.bias",
]


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(8)
    ref, ref_m = import_reference()

    # ---- 1. KAT on the real checkpoint
    sd = torch.load(os.path.join(REF, "best_ptbxl.pth"), map_location="cpu")
    x = torch.randn(4, 1, 2476, generator=torch.Generator().manual_seed(1234))
    nets = []
    for cls in (ref.ResNet1D_SE, oracle_model.ResNet1D_SE):
        net = cls(1, 2)
        net.load_state_dict(sd, strict=True)
        net.eval()
        with torch.no_grad():
            nets.append(net(x))
    assert_same(nets[0], nets[1], "ptbxl KAT")
    expect = torch.tensor([[3.9425, 0.3780], [4.0321, 0.0800], [3.9976, 0.2021], [3.8703, 0.1938]])
    assert torch.allclose(nets[1], expect, atol=1e-3), nets[1]
    shutil.copy(os.path.join(REF, "best_ptbxl.pth"), os.path.join(GOLDEN_DIR, "best_ptbxl.pth"))
    torch.save({"input_seed": 1234, "logits": nets[1]}, os.path.join(GOLDEN_DIR, "ptbxl_kat.pt"))
    print("ptbxl KAT ok:", nets[1].tolist())

    # ---- 2. fusion model, procedural weights
    ora_m = make_oracle(seed=7)
    sd = {k: v.clone() for k, v in ora_m.state_dict().items()}  # clones: state_dict() aliases live buffers
    missing, unexpected = ref_m.load_state_dict(sd, strict=True)
    assert len(sd) == 229, len(sd)
    golden = {"checksums": state_checksums(sd), "cases": {}}

    for name, (B, H, W, L) in {"small": (4, 64, 160, 600), "native": (2, 250, 2500, 2476)}.items():
        inputs = make_inputs(100 + B, B, H, W, L)
        case = {"shape": (B, H, W, L), "input_seed": 100 + B}
        # eval
        o, loss, grads = run_pair(ref_m, ora_m, inputs, "eval", 11)
        case["eval"] = {"outputs": [t.detach().clone() for t in o], "loss": loss}
        # train with dropout active (bit-identity of the oracle incl. RNG consumption order)
        run_pair(ref_m, ora_m, inputs, "train", 12)
        # restore running stats changed by the train-mode pass
        ref_m.load_state_dict(sd)
        ora_m.load_state_dict(sd)
        # train with dropout disabled: the golden the CUDA path is compared against
        set_dropout(ref_m, 0.0)
        set_dropout(ora_m, 0.0)
        o, loss, grads = run_pair(ref_m, ora_m, inputs, "train", 13)
        case["train_p0"] = {
            "outputs": [t.detach().clone() for t in o],
            "loss": loss,
            "grad_norms": {k: float(v.double().norm()) for k, v in grads.items()},
            "grads": {k: grads[k].clone() for k in SMALL_GRAD_KEYS},
            "bn_after": {k: v.clone() for k, v in ora_m.state_dict().items()
                         if ("running_" in k or "num_batches" in k) and v.numel() <= 64},
        }
        # one Adam step on both (train.py:43 defaults, lr 1e-4) -> parameters stay bit-identical
        opts = [torch.optim.Adam(m.parameters(), lr=1e-4) for m in (ref_m, ora_m)]
        for opt in opts:
            opt.step()
        for (k, a), (_, b) in zip(ref_m.named_parameters(), ora_m.named_parameters()):
            assert_same(a.detach(), b.detach(), f"param after Adam {k}")
        case["train_p0"]["param_delta_norms"] = {
            k: float((p.detach() - sd[k]).double().norm()) for k, p in ora_m.named_parameters()
        }
        golden["cases"][name] = case
        set_dropout(ref_m, 0.3)
        set_dropout(ora_m, 0.3)
        ref_m.load_state_dict(sd)
        ora_m.load_state_dict(sd)
        print(f"case {name}: oracle == reference (eval, train+dropout, train p=0, Adam step); loss={loss.item():.6f}")

    # ---- 3. FocalLoss + 12-lead signal model (cfg2 shape, small batch)
    torch.manual_seed(21)
    o_net = oracle_model.ResNet1D_SE(12, 2)  # first model after the seed: tests rebuild it the same way
    nets = [ref.ResNet1D_SE(12, 2), o_net]
    nets[0].load_state_dict(nets[1].state_dict())
    x = torch.randn(4, 12, 5000, generator=torch.Generator().manual_seed(5))
    y = torch.tensor([0, 1, 1, 0])
    outs = []
    for net in nets:
        net.train()
        set_dropout(net, 0.0)
        lo = net(x)
        loss = oracle_model.FocalLoss()(lo, y)
        loss.backward()
        outs.append((lo.detach(), loss.detach(), net.initial[0].weight.grad.clone()))
    assert_same(outs[0][0], outs[1][0], "12-lead logits")
    assert_same(outs[0][2], outs[1][2], "12-lead stem grad")
    golden["signal12"] = {"init_seed": 21, "input_seed": 5, "labels": y, "logits": outs[1][0], "focal_loss": outs[1][1],
                          "stem_grad_norm": float(outs[1][2].double().norm())}
    print("12-lead ResNet1D_SE + FocalLoss: oracle == reference; loss =", outs[1][1].item())

    torch.save(golden, os.path.join(GOLDEN_DIR, "fusion_g2.pt"))
    print("wrote", os.path.join(GOLDEN_DIR, "fusion_g2.pt"), os.path.getsize(os.path.join(GOLDEN_DIR, "fusion_g2.pt")), "bytes")


if __name__ == "__main__":
    main()

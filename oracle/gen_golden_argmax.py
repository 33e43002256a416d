"""Golden logits for the label-argmax test over many rows at the native image size (BASELINE.json north_star: "label
argmax bit-exact").  Run in the build container only (needs /root/reference):

    python oracle/gen_golden_argmax.py

The REAL reference model (multimodal_paper_modal_balance.ECGMultimodalModel) with the procedural weights of
tests/golden_util.make_oracle(7), eval mode, 256 rows of tests/golden_util.make_varied_inputs at 3x250x2500 in chunks of
8 (chunk c uses seed 9000 + c); the oracle is asserted bit-identical on the first chunk.  Stores the four logit
tensors [256, 2] and the reference margins -> tests/golden/argmax_native.pt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from golden_util import GOLDEN_DIR, make_oracle, make_varied_inputs  # noqa: E402
from gen_golden import import_reference  # noqa: E402

ROWS, CHUNK, SEED0 = 256, 8, 9000
SHAPE = (250, 2500, 2476)


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    ref, ref_m = import_reference()
    ora = make_oracle(seed=7)
    ref_m.load_state_dict({k: v.clone() for k, v in ora.state_dict().items()}, strict=True)
    ref_m.eval()
    ora.eval()
    outs = [[], [], [], []]
    with torch.no_grad():
        for c in range(ROWS // CHUNK):
            image, ecg, clin, _ = make_varied_inputs(SEED0 + c, CHUNK, *SHAPE)
            o = ref_m(image, ecg, clin)
            if c == 0:
                for a, b in zip(o[:4], ora(image, ecg, clin)[:4]):
                    assert torch.equal(a, b), "oracle != reference"
            for i in range(4):
                outs[i].append(o[i].clone())
            print(f"chunk {c + 1}/{ROWS // CHUNK}", flush=True)
    logits = [torch.cat(t) for t in outs]
    z = logits[3]
    margin = (z[:, 1] - z[:, 0]).abs()
    print("fusion logits: mean |margin| %.4f, min %.5f; class counts %s" %
          (margin.mean(), margin.min(), torch.bincount(z.argmax(1), minlength=2).tolist()))
    # at random init the class-1 bias decides nearly every row the same way: the test shifts the class-1 bias of the
    # last Linear by -median(z1 - z0) (on both sides) so that both classes occur and the margins straddle zero
    shift = -float((z[:, 1] - z[:, 0]).median())
    centred = (z[:, 1] - z[:, 0] + shift)
    print("bias shift %.6f -> class counts %s" % (shift, [int((centred <= 0).sum()), int((centred > 0).sum())]))
    for thr in (0.01, 0.02, 0.04, 0.08):
        print(f"  rows with |centred margin| > {thr}: {(centred.abs() > thr).float().mean():.3f}")
    torch.save({"rows": ROWS, "chunk": CHUNK, "seed0": SEED0, "shape": SHAPE, "logits": logits, "bias_shift": shift},
               os.path.join(GOLDEN_DIR, "argmax_native.pt"))


if __name__ == "__main__":
    main()

"""Golden vectors for the fused L1 + SSIM loss, produced by the REFERENCE'S OWN utils/loss.py (imported from
/root/reference in the build container, CPU): loss value and gradient with respect to the image for seeded inputs.

    python tests/golden/make_golden_loss.py        # writes tests/golden/loss_*.npz
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from utils.loss import l1_loss, ssim  # noqa: E402  (reference utils/loss.py:83-84, 96-135)

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
CASES = [("loss_37x53", 3, 37, 53, 11, 0.2), ("loss_64x48", 3, 64, 48, 12, 0.2), ("loss_1ch_20x9", 1, 20, 9, 13, 0.5)]


from make_golden_loss_inputs import make_inputs  # noqa: E402


if __name__ == "__main__":
    for name, C, H, W, seed, lam in CASES:
        image, gt = make_inputs(C, H, W, seed)
        image.requires_grad_(True)
        loss = (1.0 - lam) * l1_loss(image, gt) + lam * (1.0 - ssim(image, gt))  # bloomscene.py:284-287
        loss.backward()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), C=C, H=H, W=W, seed=seed, lambda_dssim=lam,
                            loss=loss.detach().numpy(), grad=image.grad.numpy())
        print(name, float(loss))

"""Seeded inputs of the loss golden vectors (shared by the generator, which needs /root/reference, and the tests)."""
import torch


def make_inputs(C, H, W, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    gt = torch.rand(C, H, W, generator=g)
    image = (gt + 0.15 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    image[:, : H // 3] = gt[:, : H // 3]  # a region where image == target exactly (sign(0) = 0 in the L1 gradient)
    return image, gt

"""Fused optimal-transport loss kernel (csrc/sinkhorn.cu) on the GPU against loss values and gradients taken through the unmodified
reference (tests/golden/sinkhorn.npz), and the loss functions with it switched on against the PyTorch formulation."""
import numpy as np
import pytest
import torch

from util import GOLDEN
from mvster_b200 import losses, train_ops

pytestmark = pytest.mark.gpu
CASES = ["d8_it3", "d8_it10", "d4_it10", "d4_it10_eps", "d8_it10_cont", "d4_it5_cont"]


@pytest.mark.parametrize("name", CASES)
def test_fused_sinkhorn_matches_reference_loss_and_gradient(name):
    z = np.load(GOLDEN / "sinkhorn.npz")
    t = {k: torch.from_numpy(z[f"{name}_{k}"]).cuda() for k in ("gt", "hypo", "attn", "mask", "loss", "grad")}
    _, _, _, _, iters, eps, cont = z[f"{name}_meta"]
    attn = t["attn"].clone().requires_grad_(True)
    assert train_ops.sinkhorn_usable(attn, int(iters))
    loss = train_ops.sinkhorn_loss(t["gt"], t["hypo"], attn, t["mask"], int(iters), float(eps), bool(cont))
    (grad,) = torch.autograd.grad(loss * 3.0, attn)
    assert abs(loss.item() - t["loss"].item()) <= 2e-5 * abs(t["loss"].item()), (loss.item(), t["loss"].item())
    assert (grad / 3.0 - t["grad"]).abs().max().item() <= 1e-4 * t["grad"].abs().max().item()


def test_loss_functions_with_the_fused_kernel_match_pytorch_ops(monkeypatch):
    """MVS4net_loss with MVSTER_TRAIN_SINKHORN=1 against the all-PyTorch formulation at a stage-like size: same total, same
    gradient into the probabilities."""
    torch.manual_seed(3)
    B, H, W = 1, 64, 80
    inputs, gts, masks = {}, {}, {}
    for s, D in enumerate((8, 8, 4, 4)):
        h, w = H >> (3 - s), W >> (3 - s)
        inv = torch.linspace(1 / 935.0, 1 / 425.0, D).reshape(1, D, 1, 1) * (1 + 0.02 * torch.rand(B, 1, h, w))
        hypo = (1.0 / inv).cuda()
        attn = torch.softmax(torch.randn(B, D, h, w) * 2, 1).cuda().requires_grad_(True)
        depth = hypo.gather(1, attn.argmax(1, keepdim=True)).squeeze(1)
        inputs[f"stage{s + 1}"] = {"hypo_depth": hypo, "attn_weight": attn, "depth": depth}
        gts[f"stage{s + 1}"] = (1.0 / (inv[:, 0] + (inv[:, -1] - inv[:, 0]) * torch.rand(B, h, w))).cuda()
        masks[f"stage{s + 1}"] = (torch.rand(B, h, w) > 0.3).float().cuda()
    kw = dict(stage_lw=[1, 1, 1, 1], l1ot_lw=[0, 1], inverse_depth=True, ot_iter=10, ot_eps=1, ot_continous=False, mono=False)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("MVSTER_TRAIN_SINKHORN", mode)
        total = losses.MVS4net_loss(inputs, gts, masks, **kw)[0]
        grads = torch.autograd.grad(total, [inputs[f"stage{s}"]["attn_weight"] for s in range(1, 5)])
        res[mode] = (total.item(), grads)
    assert abs(res["1"][0] - res["0"][0]) <= 2e-5 * abs(res["0"][0])
    for a, b in zip(res["1"][1], res["0"][1]):
        assert (a - b).abs().max().item() <= 2e-4 * b.abs().max().item()

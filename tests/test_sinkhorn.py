"""Optimal-transport loss (models/mvs4net_utils.py:1096-1142): the PyTorch restatement in mvster_b200/losses.py and the fused
forward + backward kernel (csrc/sinkhorn.cu, here as compiled source on the CPU emulation) against loss values AND autograd
gradients taken through the unmodified reference (tests/golden/sinkhorn.npz, oracle/make_sinkhorn_golden.py)."""
import numpy as np
import pytest
import torch

from test_emu_kernels import emu, emu_lib  # noqa: F401  (fixtures)
from util import GOLDEN
from mvster_b200 import capi, losses, train_ops

CASES = ["d8_it3", "d8_it10", "d4_it10", "d4_it10_eps", "d8_it10_cont", "d4_it5_cont"]


def load(name):
    z = np.load(GOLDEN / "sinkhorn.npz")
    t = {k: torch.from_numpy(z[f"{name}_{k}"]) for k in ("gt", "hypo", "attn", "mask", "loss", "grad")}
    B, D, H, W, iters, eps, cont = z[f"{name}_meta"]
    return t, int(iters), float(eps), bool(cont)


@pytest.mark.parametrize("name", CASES)
def test_torch_restatement_matches_reference_loss_and_gradient(name):
    t, iters, eps, cont = load(name)
    attn = t["attn"].clone().requires_grad_(True)
    loss = losses.sinkhorn(t["gt"], t["hypo"], attn, t["mask"], iters, eps, cont)[1]
    (grad,) = torch.autograd.grad(loss, attn)
    assert abs(loss.item() - t["loss"].item()) <= 1e-6 * abs(t["loss"].item())
    assert (grad - t["grad"]).abs().max().item() <= 1e-5 * t["grad"].abs().max().item()


@pytest.mark.parametrize("name", CASES)
def test_fused_kernel_on_cpu_matches_reference_loss_and_gradient(emu, name):
    t, iters, eps, cont = load(name)
    attn = t["attn"].clone().requires_grad_(True)
    loss = train_ops.sinkhorn_loss(t["gt"], t["hypo"], attn, t["mask"], iters, eps, cont)
    (grad,) = torch.autograd.grad(loss * 3.0, attn)  # an upstream factor, as the stage weights of MVS4net_loss apply
    assert abs(loss.item() - t["loss"].item()) <= 2e-5 * abs(t["loss"].item()), (loss.item(), t["loss"].item())
    assert (grad / 3.0 - t["grad"]).abs().max().item() <= 1e-4 * t["grad"].abs().max().item()
    assert (grad[:, :, ~t["mask"][0]] == 0).all() if t["mask"].shape[0] == 1 else True  # nothing flows through unmasked pixels


def test_fused_kernel_without_gradient_and_argument_errors(emu):
    t, iters, eps, cont = load("d4_it10")
    loss_px, grad = capi.sinkhorn(t["gt"], t["hypo"], t["attn"], t["mask"], iters, eps, cont, want_grad=False)
    assert grad is None and abs((loss_px.sum() / t["mask"].sum()).item() - t["loss"].item()) <= 2e-5 * abs(t["loss"].item())
    with pytest.raises(Exception, match="iters"):
        capi.sinkhorn(t["gt"], t["hypo"], t["attn"], t["mask"], 33, eps, cont)
    assert not train_ops.sinkhorn_usable(t["attn"], 10)       # host tensors: the product path stays on PyTorch ops
    assert train_ops.sinkhorn_enabled()                         # default on (MVSTER_TRAIN_SINKHORN=0 switches it off)


def test_loss_functions_route_through_the_fused_kernel_on_cpu(emu, monkeypatch):
    """MVS4net_loss with MVSTER_TRAIN_SINKHORN=1 (the kernel on the emulation; `usable` patched to accept host tensors) against the
    all-PyTorch formulation: same total, same gradient into every stage's probabilities."""
    torch.manual_seed(3)
    B, H, W = 1, 16, 24
    inputs, gts, masks = {}, {}, {}
    for s, D in enumerate((8, 8, 4, 4)):
        h, w = max(H >> (3 - s), 2), max(W >> (3 - s), 3)
        inv = torch.linspace(1 / 935.0, 1 / 425.0, D).reshape(1, D, 1, 1) * (1 + 0.02 * torch.rand(B, 1, h, w))
        hypo = 1.0 / inv
        attn = torch.softmax(torch.randn(B, D, h, w) * 2, 1).requires_grad_(True)
        inputs[f"stage{s + 1}"] = {"hypo_depth": hypo, "attn_weight": attn, "depth": hypo.gather(1, attn.argmax(1, keepdim=True)).squeeze(1)}
        gts[f"stage{s + 1}"] = 1.0 / (inv[:, 0] + (inv[:, -1] - inv[:, 0]) * torch.rand(B, h, w))
        masks[f"stage{s + 1}"] = (torch.rand(B, h, w) > 0.3).float()
    kw = dict(stage_lw=[1, 2, 1, 0.5], l1ot_lw=[0, 1], inverse_depth=True, ot_iter=10, ot_eps=1, ot_continous=False, mono=False)
    monkeypatch.setattr(train_ops, "sinkhorn_usable", lambda attn, iters: attn.dtype == torch.float32 and attn.shape[1] in (4, 8))
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("MVSTER_TRAIN_SINKHORN", mode)
        total = losses.MVS4net_loss(inputs, gts, masks, **kw)[0]
        res[mode] = (total.item(), torch.autograd.grad(total, [inputs[f"stage{s}"]["attn_weight"] for s in range(1, 5)]))
    assert abs(res["1"][0] - res["0"][0]) <= 2e-5 * abs(res["0"][0])
    for a, b in zip(res["1"][1], res["0"][1]):
        assert (a - b).abs().max().item() <= 2e-4 * b.abs().max().item()

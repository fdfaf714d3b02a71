"""GPU parity of the native training step (C ABI exa_train_forward /
exa_train_backward behind UNet3D.forward in train() mode) against the CPU oracle
oracle/train_ref.py, which tests/test_train_oracle.py pins to goldens of the unmodified
reference.  Tolerances are written where they are used:
  * fp32 validation mode: the whole backward graph to ~1e-4 (relative L2 per gradient tensor)
  * bf16 mode (tensor cores, bf16 activations and gradients): mixed-precision level
"""

import copy

import numpy as np
import pytest
import torch

from helpers import state_dict_for

pytestmark = pytest.mark.gpu

FP32_LOGIT_TOL = 2e-4
FP32_GRAD_REL = 2e-3     # ||g - g_ref|| / ||g_ref|| per tensor (fp32 sums in different orders)
BF16_LOGIT_TOL = 8e-2    # max |logit error|; sigmoid error stays <= 1e-2 (north_star tolerance)
BF16_GRAD_REL = 8e-2     # bf16 activations AND gradients through 18 conv + BatchNorm layers
BF16_GRAD_COS = 0.995


def _model(sd, precision):
    from aind_exaspim_neuron_segmentation_b200 import UNet3D

    m = UNet3D(output_channels=sd["outc.conv.weight"].shape[0], precision=precision)
    m.load_state_dict(sd, strict=True)
    return m.to("cuda:0")


def _step(model, x, y, grad_scale=1.0):
    model.train()
    model.zero_grad(set_to_none=True)
    logits = model(x.cuda())
    loss = torch.nn.BCEWithLogitsLoss()(logits, y.cuda())
    (loss * grad_scale).backward()
    return logits.detach().cpu(), float(loss.detach())


def _rel(a, b):
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))


def _compare_grads(model, ref, rel_tol, cos_tol=None):
    """Every parameter's gradient against the oracle.  Conv biases in front of a training-mode
    BatchNorm have a mathematically zero gradient (rounding noise on both sides): they are
    compared on the scale of the same conv's weight gradient."""
    worst = {}
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        g, r = p.grad.detach().cpu(), ref["grads"][name]
        assert g.shape == r.shape, name
        assert torch.isfinite(g).all(), name
        is_conv_bias = name.endswith(".bias") and ".double_conv." in name and \
            name.split(".")[-2] in ("0", "3")
        if is_conv_bias:
            wref = ref["grads"][name[:-4] + "weight"]
            scale = float(wref.double().norm())
            assert float(g.double().norm()) <= 1e-2 * scale + 1e-6, (name, float(g.norm()), scale)
            continue
        rel = _rel(g, r)
        worst[name] = rel
        assert rel <= rel_tol, (name, rel)
        if cos_tol is not None:
            cos = float((g.double() * r.double()).sum() /
                        (g.double().norm() * r.double().norm()).clamp_min(1e-30))
            assert cos >= cos_tol, (name, cos)
    return worst


@pytest.mark.parametrize("batch,patch", [(2, (32, 32, 32)), (3, (16, 32, 48))])
def test_train_step_fp32_mode_matches_oracle(batch, patch):
    from oracle.train_ref import train_inputs, train_step_ref

    sd = state_dict_for("rescaled", 11)
    x, y = train_inputs(21, batch, patch)
    ref = train_step_ref(x, y, sd)
    model = _model(sd, "fp32")
    logits, loss = _step(model, x, y)
    assert (logits - ref["logits"]).abs().max().item() <= FP32_LOGIT_TOL
    assert abs(loss - ref["loss"]) <= 1e-5
    worst = _compare_grads(model, ref, FP32_GRAD_REL)
    print("fp32 worst grad rel err:", max(worst.items(), key=lambda kv: kv[1]))
    # BatchNorm running statistics after the step (momentum 0.1, unbiased variance)
    for key, stat in ref["stats"].items():
        got = model.state_dict()[key].cpu()
        np.testing.assert_allclose(got.numpy(), stat.numpy(), atol=2e-5, rtol=2e-4, err_msg=key)
    for key, v in model.state_dict().items():
        if key.endswith("num_batches_tracked"):
            assert int(v) == 1, key


@pytest.mark.parametrize("batch,patch", [(2, (32, 32, 32)), (3, (16, 32, 48)), (1, (64, 64, 64))])
def test_train_step_bf16_mode_matches_oracle(batch, patch):
    from oracle.train_ref import train_inputs, train_step_ref

    sd = state_dict_for("rescaled", 12)
    x, y = train_inputs(22, batch, patch)
    ref = train_step_ref(x, y, sd)
    model = _model(sd, "bf16")
    logits, loss = _step(model, x, y)
    err = (logits - ref["logits"]).abs().max().item()
    serr = (torch.sigmoid(logits) - torch.sigmoid(ref["logits"])).abs().max().item()
    print(f"bf16 logits max err {err:.3e}, sigmoid {serr:.3e}, loss {loss:.6f} vs {ref['loss']:.6f}")
    assert err <= BF16_LOGIT_TOL and serr <= 1e-2
    assert abs(loss - ref["loss"]) <= 2e-3
    worst = _compare_grads(model, ref, BF16_GRAD_REL, BF16_GRAD_COS)
    print("bf16 worst grad rel err:", max(worst.items(), key=lambda kv: kv[1]))
    for key, stat in ref["stats"].items():
        got = model.state_dict()[key].cpu()
        np.testing.assert_allclose(got.numpy(), stat.numpy(), atol=5e-3, rtol=2e-2, err_msg=key)


def test_gradscaler_scale_passes_through_and_grads_accumulate():
    """scaler.scale(loss).backward() (train.py:140): the scale reaches every gradient; a second
    backward without zero_grad accumulates, as torch does for the reference model."""
    from oracle.train_ref import train_inputs

    sd = state_dict_for("rescaled", 13)
    x, y = train_inputs(23, 2, (32, 32, 32))
    model = _model(sd, "bf16")
    _step(model, x, y, grad_scale=1.0)
    g1 = {k: p.grad.clone() for k, p in model.named_parameters()}
    model2 = _model(sd, "bf16")
    _step(model2, x, y, grad_scale=1024.0)
    for k, p in model2.named_parameters():
        if k.endswith("weight"):
            assert _rel(p.grad / 1024.0, g1[k]) <= 2e-2, k
    # accumulation: same batch again, no zero_grad -> about twice the gradient (the running
    # statistics moved, the batch statistics did not)
    model.train()
    logits = model(x.cuda())
    torch.nn.BCEWithLogitsLoss()(logits, y.cuda()).backward()
    for k, p in model.named_parameters():
        if k.endswith("weight"):
            assert _rel(p.grad, 2 * g1[k]) <= 1e-2, k


def test_reference_style_training_loop_reduces_the_loss_and_feeds_predict():
    """Trainer.train_step on the drop-in module (train.py:136-147): AdamW + GradScaler-free bf16
    steps on one batch must reduce the loss; afterwards eval-mode forward uses the updated
    weights and running statistics."""
    from oracle.train_ref import train_inputs
    from oracle.unet_ref import unet_forward

    sd = state_dict_for("rescaled", 14)
    x, y = train_inputs(24, 2, (32, 32, 32))
    model = _model(sd, "bf16")
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    crit = torch.nn.BCEWithLogitsLoss()
    losses = []
    model.train()
    for _ in range(8):
        opt.zero_grad(set_to_none=True)
        hat_y = model(x.cuda())
        loss = crit(hat_y, y.cuda())
        loss.backward()
        opt.step()
        losses.append(float(loss))
    print("losses:", [round(v, 4) for v in losses])
    assert losses[-1] < losses[0] - 0.02, losses
    model.eval()
    with torch.no_grad():
        out = model(x.cuda()).cpu()
    new_sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = unet_forward(x, new_sd)
    assert (torch.sigmoid(out) - torch.sigmoid(ref)).abs().max().item() <= 1e-2


def test_stale_backward_and_unsupported_variants_raise():
    from aind_exaspim_neuron_segmentation_b200 import UNet3D
    from oracle.train_ref import train_inputs

    sd = state_dict_for("rescaled", 15)
    x, y = train_inputs(25, 2, (16, 16, 16))
    model = _model(sd, "bf16").train()
    a = model(x.cuda())
    b = model(x.cuda())
    b.sum().backward()
    with pytest.raises(RuntimeError, match="stale"):
        a.sum().backward()
    wide = UNet3D(output_channels=3, width_multiplier=2).to("cuda:0").train()
    with pytest.raises(NotImplementedError):
        wide(x.cuda())
    with pytest.raises(RuntimeError, match="more than one value"):
        model(x[:1].cuda())


def test_bce_with_logits_kernel_matches_torch():
    from aind_exaspim_neuron_segmentation_b200.machine_learning.training import bce_with_logits

    torch.manual_seed(3)
    logits = (torch.randn(2, 3, 24, 20, 28) * 4).cuda().requires_grad_(True)
    target = (torch.rand(2, 3, 24, 20, 28) > 0.6).float().cuda()
    ref = torch.nn.BCEWithLogitsLoss()(logits, target)
    (ref * 64.0).backward()
    loss, grad = bce_with_logits(logits.detach(), target, grad_scale=64.0)
    assert abs(float(loss) - float(ref)) <= 1e-6
    assert (grad - logits.grad).abs().max().item() <= 1e-9 + 1e-5 * logits.grad.abs().max().item()

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
# ||g - g_ref|| / ||g_ref|| per gradient tensor.  The step is ill-conditioned: gradients pass
# through LeakyReLU masks and max-pool choices, which flip with the last bits of the forward
# values -- torch's own fp32 and fp64 runs of the oracle differ by 1.6e-3 on these inputs
# (tests/test_train_oracle.py::test_conditioning_of_the_step), and so does the fp32 CUDA path.
FP32_GRAD_REL = 6e-3
# bf16 mode is compared with the oracle run that rounds the FORWARD values like the product does
# (oracle/train_ref.py, emulate_bf16): what is left is the rounding of the gradients themselves
# (bf16 dz and data gradients) and mask flips from accumulation-order differences.
# Even that run is not reproducible to better than ~0.1 in the logits (one bf16 rounding that
# falls the other way moves a batch statistic and with it everything downstream), so:
#   * the head and the last DoubleConv (up4), where few masks are involved: tight;
#   * every other tensor: magnitude and direction (the oracle's own emulate_bf16 run is 0.1-0.3
#     away from its fp32 run there on the same inputs);
#   * the two tensor-core pieces of the backward pass, exactly: operator-level tests below.
BF16_LOGIT_TOL = 0.4         # measured 0.07 .. 0.16 (logit std ~1.3; see above)
BF16_TOP_GRAD_REL = 2e-2     # measured 1e-3 .. 4.4e-3
BF16_NORM_RATIO = 0.2        # measured within 0.08
BF16_COS = 0.8               # measured >= 0.93


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


def test_train_step_foreground_mode_single_output_channel():
    """Trainer(affinity_mode=False) builds UNet3D(output_channels=1) (train.py:75-77): the head and
    its backward with one channel, fp32 validation mode against the oracle."""
    from oracle.train_ref import train_inputs, train_step_ref

    sd = state_dict_for("rescaled", 16, out_channels=1)
    x, y = train_inputs(26, 2, (32, 32, 32), out_channels=1)
    ref = train_step_ref(x, y, sd)
    model = _model(sd, "fp32")
    logits, loss = _step(model, x, y)
    assert logits.shape == (2, 1, 32, 32, 32)
    assert (logits - ref["logits"]).abs().max().item() <= FP32_LOGIT_TOL
    assert abs(loss - ref["loss"]) <= 1e-5
    _compare_grads(model, ref, FP32_GRAD_REL)


@pytest.mark.parametrize("batch,patch", [(2, (32, 32, 32)), (3, (16, 32, 48)), (1, (64, 64, 64))])
def test_train_step_bf16_mode_matches_oracle(batch, patch):
    from oracle.train_ref import train_inputs_structured, train_step_ref

    sd = state_dict_for("rescaled", 12)
    x, y = train_inputs_structured(22, batch, patch)
    ref = train_step_ref(x, y, sd)
    emu = train_step_ref(x, y, sd, emulate_bf16=True)
    model = _model(sd, "bf16")
    logits, loss = _step(model, x, y)
    err_emu = (logits - emu["logits"]).abs().max().item()
    err_ref = (logits - ref["logits"]).abs().max().item()
    print(f"bf16 logits max err vs emu {err_emu:.3e}, vs fp32 {err_ref:.3e}, "
          f"loss {loss:.6f} vs {emu['loss']:.6f} / {ref['loss']:.6f}")
    assert err_emu <= BF16_LOGIT_TOL and err_ref <= BF16_LOGIT_TOL
    assert abs(loss - emu["loss"]) <= 1e-3 and abs(loss - ref["loss"]) <= 2e-3
    report = {}
    for name, p in model.named_parameters():
        g = p.grad.detach().cpu().double()
        assert torch.isfinite(g).all(), name
        is_conv_bias = name.endswith(".bias") and name.split(".")[-2] in ("0", "3")
        if is_conv_bias:   # mathematically zero (training-mode BatchNorm follows)
            scale = float(ref["grads"][name[:-4] + "weight"].double().norm())
            assert float(g.norm()) <= 2e-2 * scale + 1e-6, name
            continue
        e, r = emu["grads"][name].double(), ref["grads"][name].double()
        rel_emu = float((g - e).norm() / e.norm())
        cos = float((g * r).sum() / (g.norm() * r.norm()))
        ratio = float(g.norm() / r.norm())
        report[name] = (rel_emu, cos, ratio)
        if name.startswith(("outc.", "up4.conv.double_conv.3", "up4.conv.double_conv.4")):
            assert rel_emu <= BF16_TOP_GRAD_REL, (name, rel_emu)
        assert abs(ratio - 1.0) <= BF16_NORM_RATIO, (name, ratio)
        assert cos >= BF16_COS, (name, cos)
    worst = min(report.items(), key=lambda kv: kv[1][1])
    print("bf16 grads: lowest cosine vs fp32 oracle", worst[0], worst[1])
    print("bf16 grads: top-layer rel err vs emulated oracle",
          {k: round(v[0], 4) for k, v in report.items() if k.startswith(("outc.", "up4.conv.double_conv.3"))})
    for key, stat in emu["stats"].items():
        got = model.state_dict()[key].cpu()
        np.testing.assert_allclose(got.numpy(), stat.numpy(), atol=1e-2, rtol=5e-2, err_msg=key)


def _ndhwc(t, dtype):
    return t.permute(0, 2, 3, 4, 1).contiguous().to(dtype).cuda()


WGRAD_SHAPES = [
    # (cin, cout, (B, D, H, W))
    (32, 32, (2, 8, 16, 24)), (64, 32, (1, 6, 10, 20)), (128, 64, (2, 8, 8, 8)),
    (512, 256, (2, 4, 4, 4)), (32, 64, (1, 16, 16, 16)), (256, 256, (3, 2, 2, 3)),
    (1, 32, (2, 8, 16, 16)), (1, 32, (1, 6, 10, 21)),
]


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("cin,cout,dims", WGRAD_SHAPES)
def test_conv_weight_grad_operator_matches_torch(precision, cin, cout, dims):
    """exa_conv3d_weight_grad (mma.sync kernel in bf16 mode) against torch.nn.grad.conv3d_weight
    in float64 on the SAME bf16-valued operands: only the fp32 accumulation order differs."""
    import ctypes

    from aind_exaspim_neuron_segmentation_b200 import _native

    lib = _native.lib()
    b, d, h, w = dims
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    gen = torch.Generator().manual_seed(cin * 1000 + cout + d)
    x = torch.randn((b, cin, d, h, w), generator=gen)
    dz = torch.randn((b, cout, d, h, w), generator=gen)
    if cin == 1:
        x_dev = x.contiguous().cuda()            # the raw float32 input of the stem
        xv = x
    else:
        x_dev = _ndhwc(x, dt)
        xv = x.to(dt).float()
    dz_dev = _ndhwc(dz, dt)
    dzv = dz.to(dt).float()
    dw = torch.empty((cout, cin, 3, 3, 3), dtype=torch.float32, device="cuda")
    code = lib.exa_conv3d_weight_grad(
        0, _native.PRECISION_BF16 if precision == "bf16" else _native.PRECISION_FP32,
        ctypes.c_void_p(x_dev.data_ptr()), ctypes.c_void_p(dz_dev.data_ptr()), b, d, h, w, cin, cout,
        ctypes.c_void_p(dw.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert code == 0, lib.exa_train_last_error(None)
    ref = torch.nn.grad.conv3d_weight(xv.double(), (cout, cin, 3, 3, 3), dzv.double(), padding=1)
    rel = _rel(dw.cpu(), ref)
    assert rel <= 2e-5, rel


DGRAD_SHAPES = [
    (32, 32, (2, 8, 16, 24)), (64, 32, (1, 6, 10, 20)), (128, 64, (2, 8, 8, 8)),
    (512, 256, (2, 4, 4, 4)), (32, 64, (1, 16, 16, 16)), (64, 128, (1, 8, 16, 16)),
    (256, 128, (3, 2, 2, 3)),
]


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("cin,cout,dims", DGRAD_SHAPES)
def test_conv_data_grad_operator_matches_torch(precision, cin, cout, dims):
    """exa_conv3d_data_grad (the tcgen05 conv kernels with flipped, transposed weights in bf16
    mode) against torch.nn.grad.conv3d_input in float64 on the same operands; the result is
    rounded to bf16 once."""
    import ctypes

    from aind_exaspim_neuron_segmentation_b200 import _native

    lib = _native.lib()
    b, d, h, w = dims
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    gen = torch.Generator().manual_seed(cin * 1000 + cout + h)
    dz = torch.randn((b, cout, d, h, w), generator=gen)
    wt = torch.randn((cout, cin, 3, 3, 3), generator=gen) / (27 * cout) ** 0.5
    dz_dev = _ndhwc(dz, dt)
    w_dev = wt.cuda()
    dx = torch.empty((b, d, h, w, cin), dtype=dt, device="cuda")
    code = lib.exa_conv3d_data_grad(
        0, _native.PRECISION_BF16 if precision == "bf16" else _native.PRECISION_FP32,
        ctypes.c_void_p(dz_dev.data_ptr()), ctypes.c_void_p(w_dev.data_ptr()), b, d, h, w, cin, cout,
        ctypes.c_void_p(dx.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert code == 0, lib.exa_train_last_error(None)
    wv = wt.to(dt).float() if precision == "bf16" else wt
    ref = torch.nn.grad.conv3d_input((b, cin, d, h, w), wv.double(), dz.to(dt).double(), padding=1)
    got = dx.float().cpu().permute(0, 4, 1, 2, 3)
    rel = _rel(got, ref)
    assert rel <= (4e-3 if precision == "bf16" else 1e-5), rel   # bf16: one rounding, 2^-9 / sqrt(3)
    assert (got.double() - ref).abs().max().item() <= (2 ** -7 if precision == "bf16" else 1e-4) * \
        ref.abs().max().item()


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
    old = unet_forward(x, sd)
    err = (torch.sigmoid(out) - torch.sigmoid(ref)).abs().max().item()
    moved = (torch.sigmoid(old) - torch.sigmoid(ref)).abs().max().item()
    # running statistics after 8 steps of 16-sample batches are rough, which amplifies the bf16
    # rounding of the eval-mode forward (2.3e-2 measured); the point here is that predict sees the
    # UPDATED weights and statistics, which moved the output an order of magnitude further
    assert err <= 5e-2 and moved >= 5 * err, (err, moved)


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


def test_weight_grad_mma_sync_fallback_matches_torch():
    """EXA_WGRAD=mma selects the warp-level mma.sync weight-gradient kernel (the first version of
    the step, kept as the A/B arm of the tcgen05 kernel): the bf16 operator cases again, in a fresh
    process because the knob is read once."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_train.py"),
                          "-q", "-x", "-k", "conv_weight_grad_operator and bf16"],
                         capture_output=True, text=True, timeout=600, cwd=root,
                         env={**os.environ, "EXA_WGRAD": "mma"})
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert " passed" in res.stdout and "failed" not in res.stdout, res.stdout[-500:]

"""GPU parity: the native U-Net forward (C ABI exa_forward) vs the CPU oracle."""

import numpy as np
import pytest
import torch

from helpers import state_dict_for

pytestmark = pytest.mark.gpu

BF16_LOGIT_TOL = 6e-2   # bf16 operands, fp32 accumulation, 18 layers (sigmoid tolerance is 1e-2)
FP32_LOGIT_TOL = 2e-4


def _oracle(x, sd, emulate=False):
    from oracle.unet_ref import unet_forward

    return unet_forward(x.cpu(), sd, emulate_bf16=emulate)


def _engine(sd, precision):
    from aind_exaspim_neuron_segmentation_b200.engine import Engine

    return Engine(sd, "cuda:0", precision)


@pytest.mark.parametrize("shape", [(1, 16, 16, 16), (3, 32, 32, 32), (2, 48, 32, 16), (1, 64, 64, 64)])
def test_forward_fp32_mode_matches_oracle(shape):
    sd = state_dict_for("rescaled", 5)
    eng = _engine(sd, "fp32")
    torch.manual_seed(1)
    x = torch.rand((shape[0], 1) + shape[1:])
    y = eng.forward(x.cuda()).cpu()
    ref = _oracle(x, sd)
    assert y.shape == ref.shape and y.dtype == torch.float32
    err = (y - ref).abs().max().item()
    assert err <= FP32_LOGIT_TOL, err
    assert (torch.sigmoid(y) - torch.sigmoid(ref)).abs().max().item() <= 1e-4


@pytest.mark.parametrize("shape", [(1, 16, 16, 16), (3, 32, 32, 32), (2, 48, 32, 16), (5, 16, 32, 48),
                                   (1, 96, 96, 96)])
def test_forward_bf16_mode_matches_oracle(shape):
    sd = state_dict_for("rescaled", 6)
    eng = _engine(sd, "bf16")
    torch.manual_seed(2)
    x = torch.rand((shape[0], 1) + shape[1:])
    y = eng.forward(x.cuda()).cpu()
    ref = _oracle(x, sd)
    emu = _oracle(x, sd, emulate=True)
    # against the bf16-emulating oracle only accumulation order / rounding ties differ
    err_emu = (y - emu).abs().max().item()
    err_ref = (torch.sigmoid(y) - torch.sigmoid(ref)).abs().max().item()
    assert err_ref <= 1e-2, (err_ref, err_emu)          # north_star tolerance (BF16)
    assert (y - ref).abs().max().item() <= BF16_LOGIT_TOL
    assert err_emu <= 3e-2, err_emu


def test_forward_default_init_and_single_channel():
    sd = state_dict_for("default", 0, out_channels=1)
    eng = _engine(sd, "bf16")
    assert eng.out_channels == 1
    x = torch.rand(2, 1, 32, 32, 32)
    y = eng.forward(x.cuda()).cpu()
    ref = _oracle(x, sd)
    assert y.shape == (2, 1, 32, 32, 32)
    assert (y - ref).abs().max().item() <= 5e-3


def test_batch_is_numerically_irrelevant():
    sd = state_dict_for("rescaled", 7)
    eng = _engine(sd, "bf16")
    x = torch.rand(4, 1, 32, 32, 32).cuda()
    whole = eng.forward(x)
    parts = torch.cat([eng.forward(x[i:i + 1]) for i in range(4)])
    assert torch.equal(whole, parts)


def test_module_forward_and_error_paths():
    from aind_exaspim_neuron_segmentation_b200 import UNet3D

    torch.manual_seed(0)
    model = UNet3D(output_channels=3).cuda().eval()
    x = torch.rand(1, 1, 32, 32, 32)
    y = model(x.cuda())
    ref = _oracle(x, {k: v.cpu() for k, v in model.state_dict().items()})
    assert (y.cpu() - ref).abs().max().item() <= 5e-3
    with pytest.raises(RuntimeError):
        model(x)                      # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        model.engine().forward(torch.rand(1, 1, 20, 32, 32).cuda())   # not a multiple of 16
    n0 = model.engine().launch_count
    model(x.cuda())
    assert model.engine().launch_count > n0


# --- model variants (unet3d.py:37,56-74,254-258) ------------------------------
VARIANTS = [(False, 1), (True, 2), (False, 2)]


@pytest.mark.parametrize("trilinear,width", VARIANTS)
def test_variant_forward_matches_oracle(trilinear, width):
    """trilinear=False (ConvTranspose3d upsampling) and width_multiplier=2 against the oracle, whose
    variant path equals the reference modules exactly (tests/test_oracle_golden.py)."""
    from oracle.unet_ref import rescaled_state_dict

    sd = rescaled_state_dict(8, 3, trilinear, width)
    torch.manual_seed(3)
    x = torch.rand(2, 1, 32, 48, 32)
    ref = _oracle(x, sd)
    y32 = _engine(sd, "fp32").forward(x.cuda()).cpu()
    assert y32.shape == ref.shape
    err32 = (y32 - ref).abs().max().item()
    assert err32 <= FP32_LOGIT_TOL, err32
    assert (torch.sigmoid(y32) - torch.sigmoid(ref)).abs().max().item() <= 1e-4
    y16 = _engine(sd, "bf16").forward(x.cuda()).cpu()
    emu = _oracle(x, sd, emulate=True)
    err_ref = (torch.sigmoid(y16) - torch.sigmoid(ref)).abs().max().item()
    err_emu = (y16 - emu).abs().max().item()
    print("variant", trilinear, width, "fp32 err", err32, "bf16 sigmoid err", err_ref, "vs emulation", err_emu)
    assert err_ref <= 1e-2, (err_ref, err_emu)
    assert err_emu <= 3e-2, err_emu


@pytest.mark.parametrize("trilinear,width", VARIANTS)
def test_variant_module_predict_matches_oracle(trilinear, width):
    """The module constructor with the reference's arguments, strict state_dict loading in both
    directions' key layout, and predict() end to end (trimmed regions, unfused head)."""
    from aind_exaspim_neuron_segmentation_b200 import UNet3D, predict
    from helpers import lightsheet_volume
    from oracle.predict_ref import predict_ref
    from oracle.unet_ref import make_forward_fn, rescaled_state_dict

    sd = rescaled_state_dict(9, 3, trilinear, width)
    model = UNet3D(output_channels=3, trilinear=trilinear, width_multiplier=width)
    assert set(model.state_dict()) == set(sd)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    vol = lightsheet_volume((48, 56, 40), 10)
    kw = dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)
    out = predict(vol, model, verbose=False, **kw)
    ref = predict_ref(vol, make_forward_fn(sd), **kw)
    err = float(np.abs(out - ref).max())
    assert out.shape == ref.shape and err <= 1e-2, err
    assert np.array_equal(out == 0, ref == 0)


def test_unsupported_widths_raise():
    from aind_exaspim_neuron_segmentation_b200 import UNet3D

    for wm in (0.5, 1.5, 8):
        with pytest.raises(NotImplementedError):
            UNet3D(output_channels=3, width_multiplier=wm)


def test_validation_forward_pass_of_the_trainer():
    """The eval-mode half of Trainer.forward_pass / validate_step (train.py:159-223): under
    no_grad and model.eval(), ``hat_y = model(x); loss = BCEWithLogitsLoss()(hat_y, y)`` runs on
    the drop-in module (logits from the engine, the criterion on them by torch).  Training-mode
    forward (batch statistics) and backward are not built and raise."""
    from aind_exaspim_neuron_segmentation_b200 import UNet3D
    from oracle.unet_ref import rescaled_state_dict

    sd = rescaled_state_dict(11, 3)
    model = UNet3D(output_channels=3)
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    torch.manual_seed(4)
    x = torch.rand(2, 1, 32, 32, 32)
    y = (torch.rand(2, 3, 32, 32, 32) > 0.7).float()
    criterion = torch.nn.BCEWithLogitsLoss()
    with torch.no_grad():
        model.eval()
        hat_y = model(x.to("cuda", dtype=torch.float))
        loss = criterion(hat_y, y.to("cuda", dtype=torch.float))
    ref_logits = _oracle(x, sd)
    ref_loss = criterion(ref_logits, y)
    assert abs(float(loss) - float(ref_loss)) <= 2e-3, (float(loss), float(ref_loss))
    assert ((hat_y.cpu() > 0) == (ref_logits > 0)).float().mean().item() >= 0.995   # compute_stats' binarisation
    model.train()
    with pytest.raises(RuntimeError):
        model(x.cuda())

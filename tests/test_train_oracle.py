"""The training-step oracle (oracle/train_ref.py) against goldens of the unmodified reference
(tests/golden/train_*.npz, made by tests/golden/make_train_golden.py): loss, logits, every
parameter gradient and the BatchNorm running statistics after the step.  CPU only."""

import os

import numpy as np
import pytest
import torch

from oracle.train_ref import is_parameter, logits_grad_ref, train_inputs, train_step_ref
from oracle.unet_ref import rescaled_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    "train_b2_p32": (11, 21, 2, (32, 32, 32)),
    "train_b3_aniso": (12, 22, 3, (16, 32, 48)),
}
LOGIT_STRIDE, GRAD_STRIDE = 97, 397


@pytest.mark.parametrize("name", sorted(CASES))
def test_training_step_oracle_matches_reference_golden(name):
    wseed, iseed, batch, patch = CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sd = rescaled_state_dict(wseed)
    x, y = train_inputs(iseed, batch, patch)
    torch.set_num_threads(os.cpu_count())
    out = train_step_ref(x, y, sd)
    assert abs(out["loss"] - float(g["loss"])) <= 1e-6
    logits = out["logits"].reshape(-1).numpy()
    np.testing.assert_allclose(logits[::LOGIT_STRIDE], g["logits_sample"], atol=2e-5, rtol=0)
    n_checked = 0
    for key, grad in out["grads"].items():
        assert is_parameter(key)
        ref_norm = float(g["gnorm/" + key])
        got = grad.double()
        # conv biases in front of a training-mode BatchNorm have a mathematically zero gradient:
        # both sides hold rounding noise there, so compare against the scale of the weights' gradient
        scale = max(ref_norm, 1e-6)
        assert abs(float(got.norm()) - ref_norm) <= 1e-3 * scale + 1e-7, key
        sample = got.reshape(-1).numpy()[::GRAD_STRIDE]
        np.testing.assert_allclose(sample, g["gsample/" + key], atol=1e-3 * scale + 1e-7, rtol=0,
                                   err_msg=key)
        n_checked += 1
    assert n_checked == 74  # 18 x (conv w, conv b, bn w, bn b) + head w, b
    for key, stat in out["stats"].items():
        np.testing.assert_allclose(stat.numpy(), g["stat/" + key], atol=1e-5, rtol=1e-5, err_msg=key)
    assert len(out["stats"]) == 36


def test_logits_grad_is_the_bce_derivative():
    x = torch.randn(2, 3, 4, 4, 4, requires_grad=True)
    y = (torch.rand(2, 3, 4, 4, 4) > 0.5).float()
    loss = torch.nn.BCEWithLogitsLoss()(x, y)
    (loss * 128.0).backward()
    np.testing.assert_allclose(logits_grad_ref(x.detach(), y, 128.0).numpy(), x.grad.numpy(),
                               atol=1e-7, rtol=1e-5)


def test_conditioning_of_the_step():
    """How sharp a gradient comparison can be: torch's fp32 and fp64 runs of the SAME oracle
    differ by ~1.6e-3 (relative L2) in the early layers -- LeakyReLU masks and max-pool choices
    flip with the last bits of the forward values -- and rounding the forward values to bf16
    (emulate_bf16) moves those gradients by tens of percent.  The GPU tolerances in
    tests/test_gpu_train.py are set from these numbers."""
    sd = rescaled_state_dict(11)
    x, y = train_inputs(21, 2, (32, 32, 32))
    torch.set_num_threads(os.cpu_count())
    r32 = train_step_ref(x, y, sd)
    r64 = train_step_ref(x, y, sd, dtype=torch.float64)
    emu = train_step_ref(x, y, sd, emulate_bf16=True)

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm())

    key = "inc.double_conv.3.weight"
    assert rel(r32["grads"]["outc.conv.weight"], r64["grads"]["outc.conv.weight"]) <= 1e-5
    assert 1e-4 <= rel(r32["grads"][key], r64["grads"][key]) <= 6e-3
    assert 5e-2 <= rel(emu["grads"][key], r32["grads"][key]) <= 0.6
    assert abs(emu["loss"] - r32["loss"]) <= 1e-3

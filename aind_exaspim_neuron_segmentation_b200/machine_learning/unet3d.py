"""3D U-Net whose forward pass runs in the native sm_100a engine.

Drop-in for reference machine_learning/unet3d.py:16-336: the module tree exists to hold
parameters under exactly the reference's ``state_dict`` names (128 entries for the
configuration ``inference.load_model`` builds, ``trilinear=True, width_multiplier=1``;
SURVEY.md 8a-1), so checkpoints written by the reference's Trainer load with
``strict=True`` and vice versa.  ``trilinear=False`` (transposed-conv upsampling) and integer
``width_multiplier`` 1..4 are supported by the engine as well.  The
arithmetic is not done by these torch modules -- ``forward`` hands the input to
``Engine.forward`` (C ABI ``exa_forward``), which folds the eval-mode BatchNorm
into the convolutions and runs the CUDA kernels.  In ``train()`` mode ``forward`` goes through
the native trainer instead (``training.train_forward``: batch-statistics BatchNorm, logits with
a ``grad_fn`` whose backward yields every parameter gradient), so the reference's
``Trainer`` (train.py:123-157) runs on this module unchanged.  On a CPU tensor ``forward``
raises: there is no fallback path.
"""

import weakref

import torch
import torch.nn as nn

from ..engine import Engine

# Native engines per module, OUTSIDE the module's __dict__: ctypes handles cannot be pickled or
# deep-copied, and a copy of a model must not share (or free) the original's device buffers.
# {module: {(precision, device): (weights fingerprint, Engine)}}
_ENGINES = weakref.WeakKeyDictionary()


def engine_for_module(module, precision, state_dict=None):
    """The native engine holding ``module``'s current weights (rebuilt when they change).

    Works for any ``nn.Module`` with the reference UNet3D ``state_dict`` layout.  Eval mode is
    checked on EVERY call: the engine folds eval-mode BatchNorm (inference.py:423)."""
    if module.training:
        raise RuntimeError("the native engine implements eval-mode BatchNorm only; "
                           "call model.eval() first (inference.py:423)")
    device = next(module.parameters()).device
    sd = state_dict if state_dict is not None else module.state_dict()
    fp = tuple((k, v.data_ptr(), v._version) for k, v in sd.items())
    per_module = _ENGINES.setdefault(module, {})
    key = (precision, str(device))
    cached = per_module.get(key)
    if cached is None or cached[0] != fp:
        if cached is not None:
            cached[1].close()
        per_module[key] = (fp, Engine(sd, device, precision))
    return per_module[key][1]

_WIDTHS = (32, 64, 128, 256, 512)


def _conv_bn_act(cin, cout):
    return [nn.Conv3d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm3d(cout),
            nn.LeakyReLU(0.01, inplace=True)]


class DoubleConv(nn.Module):
    """Parameter holder for (Conv3d 3^3 -> BatchNorm3d -> LeakyReLU) x 2, unet3d.py:108-165."""

    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid = mid_channels or out_channels
        self.double_conv = nn.Sequential(*_conv_bn_act(in_channels, mid),
                                         *_conv_bn_act(mid, out_channels))


class Down(nn.Module):
    """MaxPool3d(2) then DoubleConv, unet3d.py:168-212."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool3d(2), DoubleConv(in_channels, out_channels))


class Up(nn.Module):
    """x2 upsample (trilinear, or ConvTranspose3d(k=2, s=2) when ``trilinear=False``), concat
    [skip, upsampled], DoubleConv, unet3d.py:215-289."""

    def __init__(self, in_channels, out_channels, trilinear=True):
        super().__init__()
        if trilinear:
            self.up = nn.Upsample(scale_factor=2, mode="trilinear", align_corners=True)
            self.conv = DoubleConv(in_channels, out_channels, mid_channels=in_channels // 2)
        else:
            self.up = nn.ConvTranspose3d(in_channels, in_channels // 2, kernel_size=2, stride=2)
            self.conv = DoubleConv(in_channels, out_channels)


class OutConv(nn.Module):
    """1x1x1 head, unet3d.py:292-336."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size=1)


class UNet3D(nn.Module):
    """Same constructor and state_dict as the reference ``UNet3D`` (unet3d.py:16-105).

    ``load_model`` constructs ``trilinear=True, width_multiplier=1`` (inference.py:419-420), the
    configuration the kernels are tuned for.  ``trilinear=False`` and ``width_multiplier`` 2, 3, 4
    run on the same engine (generic tcgen05 conv kernel where the tuned ones do not fit, a SIMT
    transposed conv, the head as a kernel of its own above 32 channels).  Widths that make a
    channel count a non-multiple of 32 (fractional multipliers) raise: the tensor-core kernels
    work in 32-channel slices.  ``precision`` selects the engine arithmetic: ``"bf16"`` (tcgen05 tensor cores, fp32
    accumulation) or ``"fp32"`` (validation mode).
    """

    def __init__(self, output_channels=1, trilinear=True, width_multiplier=1, precision="bf16"):
        super().__init__()
        c = [int(w * width_multiplier) for w in _WIDTHS]    # unet3d.py:60
        if c[0] not in (32, 64, 96, 128) or c != [c[0] << k for k in range(5)]:
            raise NotImplementedError(
                f"width_multiplier={width_multiplier} gives channels {c}: the B200 engine needs "
                "32, 64, 96 or 128 channels in the first block (width_multiplier 1, 2, 3 or 4)"
            )
        factor = 2 if trilinear else 1
        self.channels = c
        self.trilinear = trilinear
        self.precision = precision
        self.inc = DoubleConv(1, c[0])
        self.down1 = Down(c[0], c[1])
        self.down2 = Down(c[1], c[2])
        self.down3 = Down(c[2], c[3])
        self.down4 = Down(c[3], c[4] // factor)
        self.up1 = Up(c[4], c[3] // factor, trilinear)
        self.up2 = Up(c[3], c[2] // factor, trilinear)
        self.up3 = Up(c[2], c[1] // factor, trilinear)
        self.up4 = Up(c[1], c[0], trilinear)
        self.outc = OutConv(c[0], output_channels)

    def engine(self, precision=None):
        """Native engine holding this module's current weights (rebuilt when they change);
        raises in training mode."""
        return engine_for_module(self, precision or self.precision)

    def forward(self, x):
        """(B,1,D,H,W) float32 -> logits (B,C,D,H,W) float32, unet3d.py:77-105."""
        if not x.is_cuda:
            raise RuntimeError("UNet3D (B200 engine) needs CUDA tensors; there is no CPU fallback")
        if self.training:
            # train.py:200-223 (Trainer.forward_pass): batch statistics, autograd-visible
            if not self.trilinear or self.channels[0] != 32:
                raise NotImplementedError("the native trainer covers the configuration the "
                                          "reference Trainer builds (train.py:77): "
                                          "trilinear=True, width_multiplier=1")
            from .training import train_forward
            return train_forward(self, x, self.precision)
        with torch.no_grad():
            return self.engine().forward(x)

"""CPU restatement of one training step of the reference -- TEST INFRASTRUCTURE ONLY.

What ``Trainer.train_step`` does per batch (REF/machine_learning/train.py:136-140, 218-222, with
``REF`` = src/aind_exaspim_neuron_segmentation): ``model.train(); hat_y = model(x);
loss = BCEWithLogitsLoss()(hat_y, y); loss.backward()``, restated with functional torch (fp32,
CPU) on a plain ``state_dict`` so that it shares no code with the reference's module tree or
the product's.  BatchNorm3d runs in training mode: batch statistics normalise, and the running
statistics move with momentum 0.1 towards the batch mean / UNBIASED batch variance
(REF/machine_learning/unet3d.py:144,147, nn.BatchNorm3d defaults).  The parameter gradients come
from torch autograd on this functional graph -- the checker of ``exa_train_forward`` /
``exa_train_backward``.

Pinned by ``tests/golden/train_*.npz``, produced by ``tests/golden/make_train_golden.py`` from the
unmodified reference module (``tests/test_train_oracle.py``).
"""

import torch
import torch.nn.functional as F

from .unet_ref import BLOCKS

PARAM_SUFFIXES = (".weight", ".bias")


def is_parameter(key):
    """state_dict entries that are nn.Parameters (everything but the BatchNorm buffers)."""
    return key.endswith(PARAM_SUFFIXES) and "running_" not in key


class _RoundBF16(torch.autograd.Function):
    """Round to bf16 in the forward pass, identity in the backward pass (straight-through): the
    forward values of the CUDA bf16 path with an unrounded fp32 backward, so that the LeakyReLU
    masks and max-pool choices agree with the product and only gradient rounding differs."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundEncodedBF16(torch.autograd.Function):
    """The raw conv output as the product stores it (training): the conv kernels'
    epilogue applies LeakyReLU(0.01) and rounds to bf16, the readers undo the LeakyReLU."""

    @staticmethod
    def forward(ctx, t):
        e = F.leaky_relu(t, 0.01).to(torch.bfloat16).to(torch.float32)
        return torch.where(e < 0, e * 100.0, e)

    @staticmethod
    def backward(ctx, g):
        return g


def _conv_bn_act_train(x, p, new_stats, prefix, conv_idx, bn_idx, emulate_bf16=False):
    """Conv3d(k=3,p=1) -> BatchNorm3d(training) -> LeakyReLU(0.01): unet3d.py:143-148."""
    w = p[f"{prefix}.{conv_idx}.weight"]
    if emulate_bf16:
        w = _RoundBF16.apply(w)
    y = F.conv3d(x, w, p[f"{prefix}.{conv_idx}.bias"], padding=1)
    if emulate_bf16:
        y = _RoundEncodedBF16.apply(y)
    rm = p[f"{prefix}.{bn_idx}.running_mean"].clone()
    rv = p[f"{prefix}.{bn_idx}.running_var"].clone()
    y = F.batch_norm(y, rm, rv, p[f"{prefix}.{bn_idx}.weight"], p[f"{prefix}.{bn_idx}.bias"],
                     training=True, momentum=0.1, eps=1e-5)
    new_stats[f"{prefix}.{bn_idx}.running_mean"] = rm
    new_stats[f"{prefix}.{bn_idx}.running_var"] = rv
    y = F.leaky_relu(y, 0.01)
    return _RoundBF16.apply(y) if emulate_bf16 else y


def _double_conv_train(x, p, new_stats, prefix, emulate_bf16=False):
    y = _conv_bn_act_train(x, p, new_stats, prefix, 0, 1, emulate_bf16)
    return _conv_bn_act_train(y, p, new_stats, prefix, 3, 4, emulate_bf16)


def unet_forward_train(x, p, new_stats, emulate_bf16=False):
    """unet3d.py:77-105 in train() mode (trilinear=True)."""
    skips = []
    h = _double_conv_train(x, p, new_stats, BLOCKS[0][0], emulate_bf16)
    skips.append(h)
    for prefix, _ in BLOCKS[1:5]:
        h = F.max_pool3d(h, 2)
        h = _double_conv_train(h, p, new_stats, prefix, emulate_bf16)
        skips.append(h)
    skips.pop()
    for prefix, _ in BLOCKS[5:]:
        up = F.interpolate(h, scale_factor=2, mode="trilinear", align_corners=True)
        if emulate_bf16:
            up = _RoundBF16.apply(up)
        h = torch.cat([skips.pop(), up], dim=1)
        h = _double_conv_train(h, p, new_stats, prefix, emulate_bf16)
    return F.conv3d(h, p["outc.conv.weight"], p["outc.conv.bias"])


def train_step_ref(x, y, sd, grad_scale=1.0, emulate_bf16=False, dtype=torch.float32):
    """One forward + backward.  x: (B,1,D,H,W), y: (B,C,D,H,W) float32; sd: state_dict.

    Returns ``dict(logits, loss, grads={name: tensor}, stats={name: tensor})``: the logits, the
    mean BCE-with-logits loss (train.py:76,222), ``grad_scale * dLoss/dparam`` for every
    parameter, and the BatchNorm running statistics after the step.

    ``emulate_bf16=True`` reproduces the forward numerics contract of the CUDA bf16 path (conv
    weights, raw conv outputs, activations and upsampled tensors rounded to bf16 once each, fp32
    accumulation and BatchNorm, stem input and head in fp32) with an unrounded backward pass.
    Gradients through LeakyReLU masks and max-pool choices are discontinuous in the forward
    values, so this -- not the fp32 run -- is what the bf16 product is compared with tightly.
    ``dtype=torch.float64`` gives the well-conditioned truth the fp32 run itself is judged by."""
    x = torch.as_tensor(x, dtype=dtype)
    y = torch.as_tensor(y, dtype=dtype)
    p = {}
    for k, v in sd.items():
        if v.dtype == torch.int64:
            continue
        t = v.detach().clone().to(dtype)
        if is_parameter(k):
            t.requires_grad_(True)
        p[k] = t
    new_stats = {}
    logits = unet_forward_train(x, p, new_stats, emulate_bf16)
    loss = F.binary_cross_entropy_with_logits(logits, y)
    (loss * grad_scale).backward()
    grads = {k: v.grad.detach() for k, v in p.items() if v.requires_grad}
    return dict(logits=logits.detach(), loss=float(loss.detach()), grads=grads,
                stats={k: v.detach() for k, v in new_stats.items()})


def logits_grad_ref(logits, y, grad_scale=1.0):
    """grad_scale * d mean-BCE / d logits = grad_scale * (sigmoid(logits) - y) / n."""
    logits = torch.as_tensor(logits, dtype=torch.float32)
    y = torch.as_tensor(y, dtype=torch.float32)
    return grad_scale * (torch.sigmoid(logits) - y) / logits.numel()


def train_inputs(seed, batch, patch, out_channels=3):
    """Seeded inputs of a training step: x ~ U(0,1) like a normalised image patch, y binary."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand((batch, 1, *patch), generator=gen)
    y = (torch.rand((batch, out_channels, *patch), generator=gen) > 0.7).float()
    return x, y


def train_inputs_structured(seed, batch, patch):
    """A learnable batch: a smooth random field, x = its noisy sigmoid (like a normalised image
    with blurred structures), y = the three affinities of its thresholded foreground (edge to
    the previous voxel along z, y, x).  With targets that depend on the input the parameter
    gradients are coherent sums instead of random walks, which is what real training steps look
    like -- and what makes a gradient comparison across precisions meaningful."""
    gen = torch.Generator().manual_seed(seed)
    noise = torch.randn((batch, 1, *patch), generator=gen)
    box = torch.ones(1, 1, 5, 5, 5) / 125.0
    s = F.conv3d(F.conv3d(noise, box, padding=2), box, padding=2)
    s = (s - s.mean()) / s.std()
    x = (torch.sigmoid(2 * s) + 0.05 * torch.randn((batch, 1, *patch), generator=gen)).clamp(0, 1)
    fg = (s > 0.3).float()
    y = torch.cat([fg * torch.roll(fg, 1, d) for d in (2, 3, 4)], dim=1)
    return x, y

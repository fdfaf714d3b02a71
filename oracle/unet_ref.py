"""CPU restatement of the reference 3D U-Net forward -- TEST INFRASTRUCTURE ONLY.

Functional torch (fp32, CPU) driven directly by a ``state_dict`` with the
reference's key layout, so it shares no code with either the reference's
``nn.Module`` tree or the product's.  ``REF`` = src/aind_exaspim_neuron_segmentation.

``emulate_bf16=True`` reproduces the *numerics contract* of the CUDA bf16 path
(DESIGN.md): BatchNorm folded into the conv in fp32, folded weights rounded to
bf16 once (including the Cin=1 stem's), every layer output rounded to bf16 once, fp32
accumulation, stem input kept to ~fp32 accuracy (bf16 hi+lo split), fp32 1x1x1 head.  It is used to separate "rounding noise" from
"bug" when the GPU result is compared with the fp32 oracle.
"""

import torch
import torch.nn.functional as F

BLOCKS = [
    # (state_dict prefix, kind)
    ("inc.double_conv", "plain"),
    ("down1.maxpool_conv.1.double_conv", "down"),
    ("down2.maxpool_conv.1.double_conv", "down"),
    ("down3.maxpool_conv.1.double_conv", "down"),
    ("down4.maxpool_conv.1.double_conv", "down"),
    ("up1.conv.double_conv", "up"),
    ("up2.conv.double_conv", "up"),
    ("up3.conv.double_conv", "up"),
    ("up4.conv.double_conv", "up"),
]


def _bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _conv_bn_act(x, sd, prefix, conv_idx, bn_idx, emulate_bf16, round_input=True):
    """Conv3d(k=3,p=1) -> BatchNorm3d(eval) -> LeakyReLU(0.01): REF/machine_learning/unet3d.py:143-148."""
    w = sd[f"{prefix}.{conv_idx}.weight"].float()
    b = sd[f"{prefix}.{conv_idx}.bias"].float()
    g = sd[f"{prefix}.{bn_idx}.weight"].float()
    beta = sd[f"{prefix}.{bn_idx}.bias"].float()
    mu = sd[f"{prefix}.{bn_idx}.running_mean"].float()
    var = sd[f"{prefix}.{bn_idx}.running_var"].float()
    if not emulate_bf16:
        y = F.conv3d(x, w, b, padding=1)
        y = F.batch_norm(y, mu, var, g, beta, training=False, eps=1e-5)
        return F.leaky_relu(y, 0.01)
    scale = (g.double() / torch.sqrt(var.double() + 1e-5))
    wf = (w.double() * scale.view(-1, 1, 1, 1, 1)).float()
    bf = ((b.double() - mu.double()) * scale + beta.double()).float()
    wf = _bf16(wf)  # every layer's folded weights are rounded once (the stem input is NOT rounded)
    y = F.leaky_relu(F.conv3d(x, wf, None, padding=1) + bf.view(1, -1, 1, 1, 1), 0.01)
    return y


def _double_conv(x, sd, prefix, emulate_bf16, last=False):
    y = _conv_bn_act(x, sd, prefix, 0, 1, emulate_bf16)
    if emulate_bf16:
        y = _bf16(y)
    y = _conv_bn_act(y, sd, prefix, 3, 4, emulate_bf16)
    if emulate_bf16 and not last:
        y = _bf16(y)
    return y


def unet_forward(x, sd, emulate_bf16=False):
    """REF/machine_learning/unet3d.py:77-105.  x: float32 (B,1,D,H,W) -> logits (B,C,D,H,W)."""
    x = torch.as_tensor(x, dtype=torch.float32)
    with torch.no_grad():
        skips = []
        h = _double_conv(x, sd, BLOCKS[0][0], emulate_bf16)
        skips.append(h)
        for prefix, _ in BLOCKS[1:5]:
            h = F.max_pool3d(h, 2)  # unet3d.py:195
            h = _double_conv(h, sd, prefix, emulate_bf16)
            skips.append(h)
        skips.pop()  # x5 is not a skip
        for i, (prefix, _) in enumerate(BLOCKS[5:]):
            # unet3d.py:248-250 (trilinear, align_corners=True) or :254-256 (ConvTranspose3d k=2 s=2
            # when the model was built with trilinear=False), and :288 (cat [skip, upsampled])
            up_w = sd.get(f"up{i + 1}.up.weight")
            if up_w is None:
                up = F.interpolate(h, scale_factor=2, mode="trilinear", align_corners=True)
            else:
                w = _bf16(up_w.float()) if emulate_bf16 else up_w.float()
                up = F.conv_transpose3d(h, w, sd[f"up{i + 1}.up.bias"].float(), stride=2)
            if emulate_bf16:
                up = _bf16(up)
            h = torch.cat([skips.pop(), up], dim=1)
            h = _double_conv(h, sd, prefix, emulate_bf16, last=(i == 3))
        # unet3d.py:318 -- the head stays fp32 in both modes
        logits = F.conv3d(h, sd["outc.conv.weight"].float(), sd["outc.conv.bias"].float())
    return logits


def make_forward_fn(sd, emulate_bf16=False):
    """numpy (B,1,P,P,P) -> numpy logits, for oracle.predict_ref.predict_ref."""
    def fn(x):
        return unet_forward(torch.from_numpy(x), sd, emulate_bf16).numpy()
    return fn


def rescaled_state_dict(seed, out_channels=3, trilinear=True, width_multiplier=1):
    """'Well-scaled' random-init weights of the reference architecture (SURVEY.md 8c caveat), for
    any constructor arguments of REF/machine_learning/unet3d.py:37-75.

    PyTorch's default init collapses activations (logits std ~0.13) and leaves BatchNorm an
    identity, which hides deep-layer and BN-folding bugs.  This builds a state_dict with the
    reference's exact key/shape layout but He-normal conv weights, randomised BN statistics
    and a head scaled so that sigmoid spans roughly (0.02, 0.98).
    """
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    c = [int(w * width_multiplier) for w in (32, 64, 128, 256, 512)]
    f = 2 if trilinear else 1

    def up(cin, cout):   # DoubleConv of an Up block: mid = in/2 with trilinear, = out otherwise
        return (cin, cin // 2 if trilinear else cout, cout)

    chans = {
        "inc.double_conv": (1, c[0], c[0]),
        "down1.maxpool_conv.1.double_conv": (c[0], c[1], c[1]),
        "down2.maxpool_conv.1.double_conv": (c[1], c[2], c[2]),
        "down3.maxpool_conv.1.double_conv": (c[2], c[3], c[3]),
        "down4.maxpool_conv.1.double_conv": (c[3], c[4] // f, c[4] // f),
        "up1.conv.double_conv": up(c[4], c[3] // f),
        "up2.conv.double_conv": up(c[3], c[2] // f),
        "up3.conv.double_conv": up(c[2], c[1] // f),
        "up4.conv.double_conv": up(c[1], c[0]),
    }
    if not trilinear:
        for i, cin in enumerate((c[4], c[3], c[2], c[1])):
            # ConvTranspose3d(k=2, s=2): every output voxel sees cin inputs once
            sd[f"up{i + 1}.up.weight"] = torch.randn((cin, cin // 2, 2, 2, 2), generator=gen) / cin ** 0.5
            sd[f"up{i + 1}.up.bias"] = torch.randn((cin // 2,), generator=gen) * 0.1
    for prefix, (cin, mid, cout) in chans.items():
        for conv_idx, bn_idx, ci, co in ((0, 1, cin, mid), (3, 4, mid, cout)):
            fan_in = ci * 27
            std = (2.0 / (1 + 0.01 ** 2)) ** 0.5 / fan_in ** 0.5
            sd[f"{prefix}.{conv_idx}.weight"] = torch.randn((co, ci, 3, 3, 3), generator=gen) * std
            sd[f"{prefix}.{conv_idx}.bias"] = torch.randn((co,), generator=gen) * 0.1
            sd[f"{prefix}.{bn_idx}.weight"] = torch.rand((co,), generator=gen) + 0.5
            sd[f"{prefix}.{bn_idx}.bias"] = torch.randn((co,), generator=gen) * 0.1
            sd[f"{prefix}.{bn_idx}.running_mean"] = torch.randn((co,), generator=gen) * 0.1
            sd[f"{prefix}.{bn_idx}.running_var"] = torch.rand((co,), generator=gen) * 1.5 + 0.5
            sd[f"{prefix}.{bn_idx}.num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)
    sd["outc.conv.weight"] = torch.randn((out_channels, c[0], 1, 1, 1), generator=gen) * (1.0 / c[0] ** 0.5)
    sd["outc.conv.bias"] = torch.randn((out_channels,), generator=gen) * 0.1
    return sd

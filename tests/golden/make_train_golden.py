"""Golden vectors of ONE TRAINING STEP, produced by the UNMODIFIED reference module.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_train_golden.py

What it runs is what ``Trainer.train_step`` runs per batch (reference
machine_learning/train.py:136-140, 218-222) without autocast: ``model.train()``,
``hat_y = model(x)``, ``loss = nn.BCEWithLogitsLoss()(hat_y, y)``, ``loss.backward()`` on the
reference's own ``UNet3D`` (CPU, fp32).  The 13 M gradient values are reduced to small
fixtures: per parameter the float64 L2 norm, the sum and a strided sample; the logits as a
strided sample plus moments; the BatchNorm running statistics after the step in full.
"""

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import import_reference  # noqa: E402
from oracle.train_ref import train_inputs  # noqa: E402
from oracle.unet_ref import rescaled_state_dict  # noqa: E402

CASES = {
    # name: (weights seed, input seed, batch, patch)
    "train_b2_p32": (11, 21, 2, (32, 32, 32)),
    "train_b3_aniso": (12, 22, 3, (16, 32, 48)),
}
LOGIT_STRIDE = 97
GRAD_STRIDE = 397


def sample(t, stride):
    flat = t.detach().reshape(-1).double().numpy()
    return flat[::stride].astype(np.float32)


def main():
    _, UNet3D = import_reference()
    torch.set_num_threads(os.cpu_count())
    for name, (wseed, iseed, batch, patch) in CASES.items():
        sd = rescaled_state_dict(wseed)
        model = UNet3D(output_channels=3)
        model.load_state_dict(sd, strict=True)
        model.train()
        x, y = train_inputs(iseed, batch, patch)
        hat_y = model(x)
        loss = torch.nn.BCEWithLogitsLoss()(hat_y, y)
        loss.backward()
        out = {"loss": np.float64(loss.item()),
               "logits_sample": sample(hat_y, LOGIT_STRIDE),
               "logits_mean": np.float64(hat_y.double().mean().item()),
               "logits_std": np.float64(hat_y.double().std().item())}
        for k, p in model.named_parameters():
            g = p.grad.double()
            out["gnorm/" + k] = np.float64(g.norm().item())
            out["gsum/" + k] = np.float64(g.sum().item())
            out["gsample/" + k] = sample(p.grad, GRAD_STRIDE)
        for k, b in model.named_buffers():
            if b.dtype == torch.int64:
                out["counter/" + k] = np.int64(b.item())
            else:
                out["stat/" + k] = b.detach().numpy().copy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "loss", loss.item(), "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

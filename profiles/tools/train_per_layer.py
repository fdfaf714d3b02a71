"""Development aid: one training step on the GPU against the CPU oracle, errors per layer.
usage: python profiles/tools/train_per_layer.py [fp32|bf16] [batch] [pz py px]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import state_dict_for  # noqa: E402
from oracle.train_ref import train_inputs, train_step_ref  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
patch = tuple(int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (32, 32, 32)

from aind_exaspim_neuron_segmentation_b200 import UNet3D  # noqa: E402

sd = state_dict_for("rescaled", 11)
x, y = train_inputs(21, batch, patch)
ref = train_step_ref(x, y, sd)
emu = train_step_ref(x, y, sd, emulate_bf16=True) if precision == "bf16" else ref
model = UNet3D(output_channels=3, precision=precision)
model.load_state_dict(sd, strict=True)
model = model.to("cuda:0").train()
logits = model(x.cuda())
loss = torch.nn.BCEWithLogitsLoss()(logits, y.cuda())
loss.backward()
torch.cuda.synchronize()
print(f"[{precision} B={batch} P={patch}] loss {float(loss):.6f} ref {ref['loss']:.6f}  "
      f"logits max err {(logits.detach().cpu() - ref['logits']).abs().max().item():.3e}  "
      f"vs emu {(logits.detach().cpu() - emu['logits']).abs().max().item():.3e}")
new = {k: v.detach().cpu() for k, v in model.state_dict().items()}


def rel(a, b):
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))


print("forward, per layer (running statistics after the step):")
for k, v in ref["stats"].items():
    if k.endswith("running_mean"):
        kv = k.replace("running_mean", "running_var")
        print(f"  {k[:-13]:42s} mean rel {rel(new[k], v):.2e}  var rel {rel(new[kv], ref['stats'][kv]):.2e}")
print("backward, last layer first:")
names = [n for n, _ in model.named_parameters()]
grads = dict(model.named_parameters())
for n in reversed(names):
    g, r = grads[n].grad.detach().cpu(), ref["grads"][n]
    print(f"  {n:48s} |ref| {float(r.norm()):.3e}  |got| {float(g.norm()):.3e}  rel {rel(g, r):.2e}  "
          f"vs emu {rel(g, emu['grads'][n]):.2e}")

"""ncu target: one warm-up training step, then one step between cudaProfilerStart/Stop
(`ncu --profile-from-start off`).  B = 4 patches of 96^3 keeps the save/restore between ncu's
replay passes short; the kernels are the ones of the B = 16 bench step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from aind_exaspim_neuron_segmentation_b200 import UNet3D  # noqa: E402
from oracle.train_ref import train_inputs_structured  # noqa: E402  (seeded inputs only)
from oracle.unet_ref import rescaled_state_dict  # noqa: E402  (weights recipe only)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = UNet3D(output_channels=3)
model.load_state_dict(rescaled_state_dict(0), strict=True)
model = model.cuda().train()
x, y = train_inputs_structured(1, B, (96, 96, 96))
x, y = x.cuda(), y.cuda()
crit = torch.nn.BCEWithLogitsLoss()


def step():
    model.zero_grad(set_to_none=True)
    loss = crit(model(x), y)
    loss.backward()
    return loss


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss.detach()))

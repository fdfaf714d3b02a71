"""Development aid: exa_conv3d_weight_grad against torch, error per tap.
usage: python profiles/tools/wgrad_per_tap.py cin cout B D H W"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from aind_exaspim_neuron_segmentation_b200 import _native  # noqa: E402

cin, cout, b, d, h, w = (int(v) for v in sys.argv[1:7])
lib = _native.lib()
gen = torch.Generator().manual_seed(1)
x = torch.randn((b, cin, d, h, w), generator=gen).bfloat16().float()
dz = torch.randn((b, cout, d, h, w), generator=gen).bfloat16().float()
xd = x.permute(0, 2, 3, 4, 1).contiguous().bfloat16().cuda()
dd = dz.permute(0, 2, 3, 4, 1).contiguous().bfloat16().cuda()
dw = torch.zeros((cout, cin, 3, 3, 3), dtype=torch.float32, device="cuda")
code = lib.exa_conv3d_weight_grad(0, 0, ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(dd.data_ptr()),
                                  b, d, h, w, cin, cout, ctypes.c_void_p(dw.data_ptr()), None)
print("code", code, lib.exa_train_last_error(None))
ref = torch.nn.grad.conv3d_weight(x.double(), (cout, cin, 3, 3, 3), dz.double(), padding=1)
got = dw.cpu().double()
print("total rel", float((got - ref).norm() / ref.norm()))
for kz in range(3):
    for ky in range(3):
        print(kz, ky, ["%.2e" % float((got[:, :, kz, ky, kx] - ref[:, :, kz, ky, kx]).norm() /
                                       ref[:, :, kz, ky, kx].norm()) for kx in range(3)],
              ["%.2e" % float(got[:, :, kz, ky, kx].norm()) for kx in range(3)])

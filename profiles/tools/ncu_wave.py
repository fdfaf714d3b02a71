"""ncu target: two predict() calls on a 160x288x288 volume = 2x4x4 windows = exactly one wave of 32
patches, so every kernel of the path appears once per layer with its production shape (the first
call warms up; read the second occurrence of each kernel in the capture).

    ncu --set full --clock-control none --import-source on -c 80 -o gpurun_out/prof_wave \
        python profiles/tools/ncu_wave.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import bench
    from aind_exaspim_neuron_segmentation_b200 import UNet3D, predict
    from oracle.unet_ref import rescaled_state_dict   # seeded weight recipe only

    model = UNet3D(output_channels=3)
    model.load_state_dict(rescaled_state_dict(0), strict=True)
    model = model.cuda().eval()
    vol = bench.synth_planes((160, 288, 288), 0, 160)
    for _ in range(2):
        out = predict(vol, model, verbose=False)
    print("ncu_wave ok", out.shape, float(out.mean()), model.engine().launch_count)


if __name__ == "__main__":
    main()

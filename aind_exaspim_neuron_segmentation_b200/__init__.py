"""B200-native drop-in for the affinity-prediction hot path of
AllenNeuralDynamics/aind-exaspim-neuron-segmentation (``inference.predict`` /
``inference.load_model``).  See DESIGN.md for the path, its boundary and the kernels.
"""

__version__ = "0.1.0"

from . import inference  # noqa: F401
from .inference import (  # noqa: F401
    affinities_to_segmentation,
    count_patches,
    generate_patch_starts,
    load_model,
    predict,
    predict_sharded,
    predict_streamed,
)
from .machine_learning.unet3d import UNet3D  # noqa: F401

"""voxelmorph.torch.layers (reference voxelmorph/torch/layers.py) -> voxelmorph_b200.layers."""
from voxelmorph_b200.layers import *          # noqa: F401,F403
import voxelmorph_b200.layers as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})

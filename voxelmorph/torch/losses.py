"""voxelmorph.torch.losses (reference voxelmorph/torch/losses.py) -> voxelmorph_b200.losses."""
from voxelmorph_b200.losses import *          # noqa: F401,F403
import voxelmorph_b200.losses as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})

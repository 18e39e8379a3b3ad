"""voxelmorph.torch.networks (reference voxelmorph/torch/networks.py) -> voxelmorph_b200.networks."""
from voxelmorph_b200.networks import *          # noqa: F401,F403
import voxelmorph_b200.networks as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})

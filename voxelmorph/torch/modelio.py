"""voxelmorph.torch.modelio (reference voxelmorph/torch/modelio.py) -> voxelmorph_b200.modelio."""
from voxelmorph_b200.modelio import *          # noqa: F401,F403
import voxelmorph_b200.modelio as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})

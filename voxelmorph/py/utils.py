"""voxelmorph.py.utils (reference voxelmorph/py/utils.py) -> voxelmorph_b200.pyutils."""
from voxelmorph_b200.pyutils import *            # noqa: F401,F403
from voxelmorph_b200.pyutils import default_unet_features, get_backend  # noqa: F401

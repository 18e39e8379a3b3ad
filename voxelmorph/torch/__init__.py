"""voxelmorph.torch (reference voxelmorph/torch/__init__.py:1-4)."""
from . import layers     # noqa: F401
from . import networks   # noqa: F401
from . import losses     # noqa: F401
from . import modelio    # noqa: F401
from . import utils      # noqa: F401

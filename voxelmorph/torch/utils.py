"""voxelmorph.torch.utils — the reference module is empty apart from imports (voxelmorph/torch/utils.py)."""

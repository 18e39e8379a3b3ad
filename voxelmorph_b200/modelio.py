"""Checkpoint surface of the reference (voxelmorph/torch/modelio.py): `store_config_args`
records constructor arguments in `self.config`; `LoadableModel.save/load` round-trip
`{'config': ..., 'model_state': ...}` through torch.save / torch.load.  Files written by the
reference load here and vice versa (keys ending in `.grid` are neither written nor required).
"""
import functools
import inspect

import torch
import torch.nn as nn


def store_config_args(func):
    """Decorator for `__init__`: saves every argument (defaults included) into `self.config`."""
    sig = inspect.signature(func)
    names = list(sig.parameters)[1:]  # drop self

    @functools.wraps(func)
    def wrapper(self, *args, **kwargs):
        cfg = {n: p.default for n, p in sig.parameters.items()
               if n != 'self' and p.default is not inspect.Parameter.empty}
        cfg.update(zip(names, args))
        cfg.update(kwargs)
        self.config = cfg
        return func(self, *args, **kwargs)

    return wrapper


class LoadableModel(nn.Module):
    """nn.Module whose architecture can be rebuilt from a checkpoint (reference modelio.py:38-77)."""

    def __init__(self, *args, **kwargs):
        if not hasattr(self, 'config'):
            raise RuntimeError('models that inherit from LoadableModel must decorate the '
                               'constructor with @store_config_args')
        super().__init__(*args, **kwargs)

    def save(self, path):
        """Write {'config', 'model_state'}; `.grid` buffers never enter the file."""
        state = {k: v for k, v in self.state_dict().items() if not k.endswith('.grid')}
        torch.save({'config': self.config, 'model_state': state}, path)

    @classmethod
    def load(cls, path, device):
        """Rebuild the model from a checkpoint written by this class or by the reference."""
        checkpoint = torch.load(path, map_location=torch.device(device))
        model = cls(**checkpoint['config'])
        model.load_state_dict(checkpoint['model_state'], strict=False)
        return model

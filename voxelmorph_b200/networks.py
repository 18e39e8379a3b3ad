"""Unet / VxmDense / ConvBlock with the reference's surface
(reference voxelmorph/torch/networks.py): same class names, constructor arguments, attributes
(`unet_model`, `flow`, `resize`, `fullsize`, `integrate`, `transformer`, `bidir`, `config`,
`Unet.final_nf`), forward signatures / return tuples and `state_dict` keys — so reference
checkpoints load unchanged — with every operator running in sm_100a kernels.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.distributions.normal import Normal

from . import layers, ops
from .modelio import LoadableModel, store_config_args


def default_unet_features():
    """reference voxelmorph/py/utils.py:16-21"""
    return [[16, 32, 32, 32], [32, 32, 32, 32, 32, 16, 16]]


def _conv_cls(ndims):
    if ndims == 2:
        return _Conv2dK3
    if ndims == 3:
        return _Conv3dK3
    raise NotImplementedError("voxelmorph_b200: %d-D convolutions are not supported (2-D and 3-D only; the "
                              "reference's SpatialTransformer handles 2-D/3-D only as well, layers.py:41-46)" % ndims)


class _Conv3dK3(nn.Conv3d):
    """nn.Conv3d parameters (weight (Cout,Cin,3,3,3), bias, default init) + the vxm conv kernel."""

    def forward(self, x):
        _check_k3(self)
        return ops.conv_k3(x, self.weight, self.bias, None)


class _Conv2dK3(nn.Conv2d):
    def forward(self, x):
        _check_k3(self)
        return ops.conv_k3(x, self.weight, self.bias, None)


def _check_k3(m):
    nd = len(m.kernel_size)
    if tuple(m.kernel_size) != (3,) * nd or tuple(m.stride) != (1,) * nd or tuple(m.padding) != (1,) * nd \
            or tuple(m.dilation) != (1,) * nd or m.groups != 1:
        raise NotImplementedError("voxelmorph_b200 conv kernels implement kernel 3, stride 1, padding 1 only")


class ConvBlock(nn.Module):
    """Convolution followed by LeakyReLU(0.2) (reference networks.py:290-305), fused in one kernel."""

    def __init__(self, ndims, in_channels, out_channels, stride=1):
        super().__init__()
        self.main = _conv_cls(ndims)(in_channels, out_channels, 3, stride, 1)
        self.activation = nn.LeakyReLU(0.2)

    def forward(self, x):
        _check_k3(self.main)
        return ops.conv_k3(x, self.main.weight, self.main.bias, self.activation.negative_slope)


class _Pool2(nn.Module):
    def forward(self, x):
        return ops.maxpool2(x)


class Unet(nn.Module):
    """
    A unet architecture (reference networks.py:12-144). Layer features can be specified directly as a
    list of encoder and decoder features or as a single integer along with a number of unet levels.
    Default features: encoder [16, 32, 32, 32], decoder [32, 32, 32, 32, 32, 16, 16].
    """

    def __init__(self, inshape=None, infeats=None, nb_features=None, nb_levels=None, max_pool=2,
                 feat_mult=1, nb_conv_per_level=1, half_res=False):
        super().__init__()
        ndims = len(inshape)
        assert ndims in [1, 2, 3], 'ndims should be one of 1, 2, or 3. found: %d' % ndims
        self.half_res = half_res

        if nb_features is None:
            nb_features = default_unet_features()
        if isinstance(nb_features, int):
            if nb_levels is None:
                raise ValueError('must provide unet nb_levels if nb_features is an integer')
            feats = np.round(nb_features * feat_mult ** np.arange(nb_levels)).astype(int)
            nb_features = [np.repeat(feats[:-1], nb_conv_per_level), np.repeat(np.flip(feats), nb_conv_per_level)]
        elif nb_levels is not None:
            raise ValueError('cannot use nb_levels if nb_features is not an integer')

        enc_nf, dec_nf = nb_features
        nb_dec_convs = len(enc_nf)
        final_convs = dec_nf[nb_dec_convs:]
        dec_nf = dec_nf[:nb_dec_convs]
        self.nb_levels = int(nb_dec_convs / nb_conv_per_level) + 1

        if isinstance(max_pool, int):
            max_pool = [max_pool] * self.nb_levels
        if any(int(s) != 2 for s in max_pool):
            raise NotImplementedError("voxelmorph_b200: only max_pool=2 is implemented")
        self.pooling = [_Pool2() for _ in max_pool]
        self.upsampling = [nn.Upsample(scale_factor=s, mode='nearest') for s in max_pool]  # attribute parity only

        prev_nf = infeats
        encoder_nfs = [prev_nf]
        self.encoder = nn.ModuleList()
        for level in range(self.nb_levels - 1):
            convs = nn.ModuleList()
            for conv in range(nb_conv_per_level):
                nf = int(enc_nf[level * nb_conv_per_level + conv])
                convs.append(ConvBlock(ndims, prev_nf, nf))
                prev_nf = nf
            self.encoder.append(convs)
            encoder_nfs.append(prev_nf)

        encoder_nfs = np.flip(encoder_nfs)
        self.decoder = nn.ModuleList()
        for level in range(self.nb_levels - 1):
            convs = nn.ModuleList()
            for conv in range(nb_conv_per_level):
                nf = int(dec_nf[level * nb_conv_per_level + conv])
                convs.append(ConvBlock(ndims, prev_nf, nf))
                prev_nf = nf
            self.decoder.append(convs)
            if not half_res or level < (self.nb_levels - 2):
                prev_nf += int(encoder_nfs[level])

        self.remaining = nn.ModuleList()
        for nf in final_convs:
            self.remaining.append(ConvBlock(ndims, prev_nf, int(nf)))
            prev_nf = int(nf)
        self.final_nf = prev_nf

    def forward(self, x):
        skips = [x]
        for level, convs in enumerate(self.encoder):
            for conv in convs:
                x = conv(x)
            skips.append(x)
            x = self.pooling[level](x)
        for level, convs in enumerate(self.decoder):
            for conv in convs:
                x = conv(x)
            if not self.half_res or level < (self.nb_levels - 2):
                x = ops.upsample2_cat(x, skips.pop())   # nearest x2 + concat fused (networks.py:137-138)
        for conv in self.remaining:
            x = conv(x)
        return x


class VxmDense(LoadableModel):
    """VoxelMorph network for (unsupervised) nonlinear registration between two images
    (reference networks.py:147-287)."""

    @store_config_args
    def __init__(self, inshape, nb_unet_features=None, nb_unet_levels=None, unet_feat_mult=1,
                 nb_unet_conv_per_level=1, int_steps=7, int_downsize=2, bidir=False, use_probs=False,
                 src_feats=1, trg_feats=1, unet_half_res=False):
        super().__init__()
        self.training = True
        object.__setattr__(self, "_dp", None)            # dist.TransparentDP once attached (not a submodule / buffer)
        object.__setattr__(self, "_dp_checked", False)
        ndims = len(inshape)
        assert ndims in [1, 2, 3], 'ndims should be one of 1, 2, or 3. found: %d' % ndims

        self.unet_model = Unet(inshape, infeats=(src_feats + trg_feats), nb_features=nb_unet_features,
                               nb_levels=nb_unet_levels, feat_mult=unet_feat_mult,
                               nb_conv_per_level=nb_unet_conv_per_level, half_res=unet_half_res)

        self.flow = _conv_cls(ndims)(self.unet_model.final_nf, ndims, kernel_size=3, padding=1)
        self.flow.weight = nn.Parameter(Normal(0, 1e-5).sample(self.flow.weight.shape))
        self.flow.bias = nn.Parameter(torch.zeros(self.flow.bias.shape))

        if use_probs:
            raise NotImplementedError('Flow variance has not been implemented in pytorch - set use_probs to False')

        if not unet_half_res and int_steps > 0 and int_downsize > 1:
            self.resize = layers.ResizeTransform(int_downsize, ndims)
        else:
            self.resize = None
        if int_steps > 0 and int_downsize > 1:
            self.fullsize = layers.ResizeTransform(1 / int_downsize, ndims)
        else:
            self.fullsize = None

        self.bidir = bidir
        down_shape = [int(dim / int_downsize) for dim in inshape]
        self.integrate = layers.VecInt(down_shape, int_steps) if int_steps > 0 else None
        self.transformer = layers.SpatialTransformer(inshape)

    def _maybe_attach_dp(self):
        """Under torchrun the model becomes data parallel by itself on its first forward (dist.TransparentDP): the
        unmodified training loop then needs no DataParallel / DistributedDataParallel wrapper."""
        if not self._dp_checked:
            object.__setattr__(self, "_dp_checked", True)
            if self.training:
                from . import dist as vdist
                object.__setattr__(self, "_dp", vdist.attach_if_distributed(self))
        return self._dp

    def save(self, path):
        if self._dp is not None and not self._dp.is_writer():
            return                       # one checkpoint per job: rank 0 writes
        super().save(path)

    def forward(self, source, target, registration=False):
        self._maybe_attach_dp()
        engine = ops.resolve_engine(self)
        if engine in ('bf16', 'bf16x3'):
            # tensor-core engine: Unet + flow head as one hand-written forward/backward (engine_bf16.py)
            from . import engine_bf16
            flow_field = engine_bf16.unet_flow(self, source, target, split=(engine == 'bf16x3'))
        else:
            x = ops.upsample_free_cat(source, target)
            x = self.unet_model(x)
            flow_field = self.flow(x)

        pos_flow = flow_field
        if self.resize:
            pos_flow = self.resize(pos_flow)
        preint_flow = pos_flow
        neg_flow = -pos_flow if self.bidir else None

        if self.integrate:
            pos_flow = self.integrate(pos_flow)
            neg_flow = self.integrate(neg_flow) if self.bidir else None
            if self.fullsize:
                pos_flow = self.fullsize(pos_flow)
                neg_flow = self.fullsize(neg_flow) if self.bidir else None

        y_source = self.transformer(source, pos_flow)
        y_target = self.transformer(target, neg_flow) if self.bidir else None

        if not registration:
            return (y_source, y_target, preint_flow) if self.bidir else (y_source, preint_flow)
        return y_source, pos_flow

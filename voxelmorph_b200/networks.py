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
        self.registration_no_grad = True     # eval-mode registration calls run without autograd bookkeeping (see forward)
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

    def flows(self, source, target):
        """(pos_flow, neg_flow, preint_flow): the U-Net's field brought to the integration resolution (`preint_flow`,
        what train.py regularises), integrated and brought back to full resolution (`pos_flow`, what warps the moving
        image; `neg_flow` its inverse when bidir).  reference networks.py:253-276."""
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
        return pos_flow, neg_flow, preint_flow

    def forward(self, source, target, registration=False):
        self._maybe_attach_dp()
        if registration and not self.training and self.registration_no_grad and torch.is_grad_enabled():
            # inference (register.py:80-87 calls model.eval() but never torch.no_grad()): nothing downstream of a
            # registration call differentiates, so no activation / VecInt state is kept ("next" row N3)
            with torch.no_grad():
                return self.forward(source, target, registration=True)
        pos_flow, neg_flow, preint_flow = self.flows(source, target)
        y_source = self.transformer(source, pos_flow)
        y_target = self.transformer(target, neg_flow) if self.bidir else None

        if not registration:
            return (y_source, y_target, preint_flow) if self.bidir else (y_source, preint_flow)
        return y_source, pos_flow


class VxmDenseSemiSupervisedSeg(LoadableModel):
    """VoxelMorph network for semi-supervised registration with segmentations ("next" row N2; BASELINE config 5).

    The torch backend of the reference has no such class; the semantics are those of its TensorFlow model
    (voxelmorph/tf/networks.py:287-366) restated over the torch layers, as SURVEY.md section 8(a) A12 specifies: the
    full-resolution `pos_flow` of the inner VxmDense is rescaled to the segmentation resolution
    (`ResizeTransform(seg_resolution)`, i.e. RescaleTransform(1 / seg_resolution)), the probabilistic (one-hot) source
    segmentation is warped LINEARLY with it, and the result joins the outputs so that a Dice loss can be attached:

        y_source, preint_flow, y_seg = model(source, target, seg_source)
        loss = image_loss(target, y_source) + w_grad * Grad(preint_flow) + w_seg * Dice(seg_target, y_seg)

    `bidir_labels` additionally warps the target segmentation with `neg_flow` (and implies bidir).  The inner network is
    `self.vxm_model` (checkpoint keys `vxm_model.*`); `kwargs` are forwarded to it."""

    @store_config_args
    def __init__(self, inshape, nb_labels, nb_unet_features=None, seg_resolution=2, bidir=False, bidir_labels=False, **kwargs):
        super().__init__()
        if bidir_labels:
            bidir = True
        ndims = len(inshape)
        self.nb_labels = int(nb_labels)
        self.bidir_labels = bool(bidir_labels)
        self.vxm_model = VxmDense(inshape, nb_unet_features=nb_unet_features, bidir=bidir, **kwargs)
        self.seg_resize = layers.ResizeTransform(seg_resolution, ndims) if seg_resolution != 1 else None
        inshape_ds = [int(d / seg_resolution) for d in inshape]
        self.seg_transformer = layers.SpatialTransformer(inshape_ds)        # linear: the segmentation is a probability map

    def forward(self, source, target, seg_source, seg_target=None, registration=False):
        vm = self.vxm_model
        vm._maybe_attach_dp()
        pos_flow, neg_flow, preint_flow = vm.flows(source, target)
        y_source = vm.transformer(source, pos_flow)
        y_target = vm.transformer(target, neg_flow) if vm.bidir else None
        seg_flow = self.seg_resize(pos_flow) if self.seg_resize else pos_flow
        y_seg = self.seg_transformer(seg_source, seg_flow)
        outs = [y_source] + ([y_target] if vm.bidir else []) + [pos_flow if registration else preint_flow, y_seg]
        if self.bidir_labels:
            if seg_target is None:
                raise ValueError("bidir_labels=True needs the target segmentation")
            nseg_flow = self.seg_resize(neg_flow) if self.seg_resize else neg_flow
            outs.append(self.seg_transformer(seg_target, nseg_flow))
        return tuple(outs)

    def save(self, path):
        dp = self.vxm_model._dp
        if dp is not None and not dp.is_writer():
            return
        super().save(path)

    def register(self, source, target):
        """The transform from source to target (full resolution), like tf get_registration_model / register."""
        with torch.no_grad():
            return self.vxm_model.flows(source, target)[0]

    def apply_transform(self, source, target, img, interp_method='linear'):
        """Predict the transform from source to target and apply it to `img` ('linear' or 'nearest')."""
        mode = 'bilinear' if interp_method == 'linear' else interp_method
        flow = self.register(source, target)
        return layers.SpatialTransformer(tuple(img.shape[2:]), mode=mode)(img, flow)

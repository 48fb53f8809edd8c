"""Drop-in `Generic_UNet` for the BraTS-2021 nnU-Net models (reference: model_architecture/generic_UNet.py).

Same constructor signature (:188-198), same attributes (`do_ds`, `conv_op`, `num_classes`, `inference_apply_nonlin`,
`input_shape_must_be_divisible_by`, ...) and — because the module tree uses the same attribute names — the same
`state_dict` keys, so nnU-Net v1 checkpoints (`model_final_checkpoint.model`) load unchanged.  The torch modules here
only HOLD the parameters: `forward` and `predict_3D` execute on the B200 engine (hand-written sm_100a kernels), not
through torch.nn.  Scope: the 3-D configuration the BraTS trainers use (Conv3d, conv pooling + transposed-conv
upsampling, no axial attention); anything else raises NotImplementedError.
"""
from copy import deepcopy

import numpy as np
import torch
from torch import nn

from .neural_network import SegmentationNetwork


def softmax_helper(x):
    return torch.softmax(x, 1)


class InitWeights_He(object):
    """nnU-Net v1 initialiser: kaiming_normal_(a=neg_slope) on (transposed) conv weights, zero biases."""

    def __init__(self, neg_slope=1e-2):
        self.neg_slope = neg_slope

    def __call__(self, module):
        if isinstance(module, (nn.Conv3d, nn.Conv2d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
            module.weight = nn.init.kaiming_normal_(module.weight, a=self.neg_slope)
            if module.bias is not None:
                module.bias = nn.init.constant_(module.bias, 0)


def _default(kwargs, fallback):
    return dict(fallback) if kwargs is None else kwargs


_NONLIN = {"negative_slope": 1e-2, "inplace": True}
_DROPOUT = {"p": 0.5, "inplace": True}
_NORM = {"eps": 1e-5, "affine": True, "momentum": 0.1}
_CONV = {"kernel_size": 3, "stride": 1, "padding": 1, "dilation": 1, "bias": True}


class ConvDropoutNormNonlin(nn.Module):
    """Parameter holder for conv -> (dropout) -> norm -> nonlin (reference :27-72).  Attribute names `conv`,
    `dropout`, `instnorm`, `lrelu` are part of the checkpoint key layout."""

    def __init__(self, input_channels, output_channels, conv_op=nn.Conv2d, conv_kwargs=None, norm_op=nn.BatchNorm2d,
                 norm_op_kwargs=None, dropout_op=nn.Dropout2d, dropout_op_kwargs=None, nonlin=nn.LeakyReLU,
                 nonlin_kwargs=None):
        super().__init__()
        self.nonlin_kwargs = _default(nonlin_kwargs, _NONLIN)
        self.dropout_op_kwargs = _default(dropout_op_kwargs, _DROPOUT)
        self.norm_op_kwargs = _default(norm_op_kwargs, _NORM)
        self.conv_kwargs = _default(conv_kwargs, _CONV)
        self.nonlin, self.dropout_op, self.conv_op, self.norm_op = nonlin, dropout_op, conv_op, norm_op
        self.conv = conv_op(input_channels, output_channels, **self.conv_kwargs)
        p = self.dropout_op_kwargs.get("p") if dropout_op is not None else None
        self.dropout = dropout_op(**self.dropout_op_kwargs) if p is not None and p > 0 else None
        if norm_op == nn.GroupNorm:
            self.instnorm = norm_op(num_channels=output_channels, **self.norm_op_kwargs)
        else:
            self.instnorm = norm_op(output_channels, **self.norm_op_kwargs)
        self.lrelu = nonlin(**self.nonlin_kwargs)

    def forward(self, x):
        raise RuntimeError("brainseg_b200 blocks hold parameters only; run the network through Generic_UNet")


class StackedConvLayers(nn.Module):
    """`num_convs` blocks, the first one optionally strided (reference :83-146); attribute `blocks`."""

    def __init__(self, input_feature_channels, output_feature_channels, num_convs, conv_op=nn.Conv2d, conv_kwargs=None,
                 norm_op=nn.BatchNorm2d, norm_op_kwargs=None, dropout_op=nn.Dropout2d, dropout_op_kwargs=None,
                 nonlin=nn.LeakyReLU, nonlin_kwargs=None, first_stride=None, basic_block=ConvDropoutNormNonlin):
        super().__init__()
        self.input_channels, self.output_channels = input_feature_channels, output_feature_channels
        conv_kwargs = _default(conv_kwargs, _CONV)
        first_kwargs = conv_kwargs
        if first_stride is not None:
            first_kwargs = deepcopy(conv_kwargs)
            first_kwargs["stride"] = first_stride
        common = (norm_op, _default(norm_op_kwargs, _NORM), dropout_op, _default(dropout_op_kwargs, _DROPOUT), nonlin,
                  _default(nonlin_kwargs, _NONLIN))
        layers = [basic_block(input_feature_channels, output_feature_channels, conv_op, first_kwargs, *common)]
        layers += [basic_block(output_feature_channels, output_feature_channels, conv_op, conv_kwargs, *common)
                   for _ in range(num_convs - 1)]
        self.blocks = nn.Sequential(*layers)


class Generic_UNet(SegmentationNetwork):
    DEFAULT_BATCH_SIZE_3D = 2
    DEFAULT_PATCH_SIZE_3D = (64, 192, 160)
    SPACING_FACTOR_BETWEEN_STAGES = 2
    BASE_NUM_FEATURES_3D = 30
    MAX_NUMPOOL_3D = 999
    MAX_NUM_FILTERS_3D = 320

    def __init__(self, input_channels, base_num_features, num_classes, num_pool, num_conv_per_stage=2,
                 feat_map_mul_on_downscale=2, conv_op=nn.Conv2d, norm_op=nn.BatchNorm2d, norm_op_kwargs=None,
                 dropout_op=nn.Dropout2d, dropout_op_kwargs=None, nonlin=nn.LeakyReLU, nonlin_kwargs=None,
                 deep_supervision=True, dropout_in_localization=False, final_nonlin=softmax_helper,
                 weightInitializer=InitWeights_He(1e-2), pool_op_kernel_sizes=None, conv_kernel_sizes=None,
                 upscale_logits=False, convolutional_pooling=False, convolutional_upsampling=False,
                 max_num_features=None, basic_block=ConvDropoutNormNonlin, seg_output_use_bias=False,
                 encoder_scale=1, axial_attention=False, heads=8, dim_heads=32, volume_shape=(128, 128, 128),
                 no_attention=[0], dropout_level=[4]):
        super().__init__()
        if conv_op != nn.Conv3d:
            raise NotImplementedError("brainseg_b200 covers the 3-D (nn.Conv3d) BraTS configuration only")
        if not (convolutional_pooling and convolutional_upsampling):
            raise NotImplementedError("only convolutional pooling + transposed-conv upsampling (nnU-Net V2 trainers)")
        if axial_attention or upscale_logits:
            raise NotImplementedError("axial attention / upscale_logits are outside the BraTS hot path")
        self.convolutional_upsampling, self.convolutional_pooling = True, True
        self.upscale_logits = False
        self.nonlin, self.nonlin_kwargs = nonlin, _default(nonlin_kwargs, _NONLIN)
        self.dropout_op, self.dropout_op_kwargs = dropout_op, _default(dropout_op_kwargs, _DROPOUT)
        self.norm_op, self.norm_op_kwargs = norm_op, _default(norm_op_kwargs, _NORM)
        self.conv_op, self.weightInitializer = conv_op, weightInitializer
        self.num_classes, self.final_nonlin = num_classes, final_nonlin
        self._deep_supervision = self.do_ds = deep_supervision
        self.do_attention = False
        self.volume_shape, self.no_attention, self.dropout_level = np.array(volume_shape), no_attention, dropout_level
        pools = [tuple(k) for k in (pool_op_kernel_sizes or [(2, 2, 2)] * num_pool)]
        kernels = [tuple(k) for k in (conv_kernel_sizes or [(3, 3, 3)] * (num_pool + 1))]
        if any(k != (3, 3, 3) for k in kernels) or any(p != (2, 2, 2) for p in pools):
            raise NotImplementedError("the sm_100a kernels implement 3x3x3 convs with 2x2x2 pooling strides")
        self.pool_op_kernel_sizes, self.conv_kernel_sizes = pools, kernels
        self.input_shape_must_be_divisible_by = np.prod(pools, 0, dtype=np.int64)
        self.conv_pad_sizes = [[1, 1, 1] for _ in kernels]
        self.max_num_features = self.MAX_NUM_FILTERS_3D if max_num_features is None else max_num_features
        self.conv_kwargs = {"stride": 1, "dilation": 1, "bias": True, "kernel_size": (3, 3, 3), "padding": [1, 1, 1]}

        def stack(cin, cout, n, stride, drop_kwargs):
            return StackedConvLayers(cin, cout, n, conv_op, self.conv_kwargs, norm_op, self.norm_op_kwargs, dropout_op,
                                     drop_kwargs, nonlin, self.nonlin_kwargs, stride, basic_block=basic_block)

        # ---- encoder: stage d > 0 carries the pooling stride in its first conv
        context, widths = [], []
        cin, cout = input_channels, base_num_features * encoder_scale
        for d in range(num_pool):
            drop = dict(self.dropout_op_kwargs)
            if d not in self.dropout_level:
                drop["p"] = 0.0
            context.append(stack(cin, cout, num_conv_per_stage, pools[d - 1] if d else None, drop))
            widths.append(cout)
            cin, cout = cout, min(int(np.round(cout * feat_map_mul_on_downscale)), self.max_num_features)
        # ---- bottleneck: (n-1) convs at `cout`, then one conv to `final`
        final = cout
        context.append(nn.Sequential(stack(cin, cout, num_conv_per_stage - 1, pools[-1], self.dropout_op_kwargs),
                                     stack(cout, final, 1, None, self.dropout_op_kwargs)))
        # ---- decoder
        loc_drop = dict(self.dropout_op_kwargs)
        if not dropout_in_localization:
            loc_drop["p"] = 0.0
        localization, tu, seg = [], [], []
        for u in range(num_pool):
            from_down = final if u == 0 else int(final / encoder_scale)
            skip = widths[-(1 + u)]
            tu.append(nn.ConvTranspose3d(from_down, skip, pools[-(u + 1)], pools[-(u + 1)], bias=False))
            final = skip
            out = int(final / encoder_scale)
            localization.append(nn.Sequential(stack(skip * 2, skip, num_conv_per_stage - 1, None, loc_drop),
                                              stack(skip, out, 1, None, loc_drop)))
            seg.append(conv_op(out, num_classes, 1, 1, 0, 1, 1, seg_output_use_bias))
        self.conv_blocks_localization = nn.ModuleList(localization)
        self.conv_blocks_context = nn.ModuleList(context)
        self.td = nn.ModuleList([])
        self.tu = nn.ModuleList(tu)
        self.seg_outputs = nn.ModuleList(seg)
        self.upscale_logits_ops = [(lambda x: x) for _ in range(num_pool - 1)]
        if weightInitializer is not None:
            self.apply(weightInitializer)

    def forward(self, x):
        """(N, C, D, H, W) float tensor -> final-resolution logits after `final_nonlin` (reference :423-446 with
        do_ds False).  Runs on the sm_100a engine; the deep-supervision tuple (training only) is out of scope."""
        if self._deep_supervision and self.do_ds:
            raise NotImplementedError("deep-supervision outputs are a training feature; set network.do_ds = False")
        logits = self.engine_for(tuple(x.shape[2:]), x.shape[0]).forward_logits(x)
        return self.final_nonlin(logits)

"""Shared builders for the UNet parity tests."""
import torch
from torch import nn


def build_dropin_unet(variant="bn", base=32, num_pool=5, in_ch=4, num_classes=3, seed=1, groups=8, encoder_scale=1,
                      max_num_features=None, nonlin="sigmoid"):
    """The drop-in Generic_UNet built the way the BraTS-2021 V2 trainers build the reference class."""
    from brainseg_b200 import generic_UNet as G
    from oracle.ref_import import randomize_norm_params

    norm_op = {"bn": nn.BatchNorm3d, "in": nn.InstanceNorm3d, "gn": nn.GroupNorm}[variant]
    norm_kwargs = {"eps": 1e-5, "affine": True}
    if variant == "gn":
        norm_kwargs["num_groups"] = groups
    torch.manual_seed(seed)
    net = G.Generic_UNet(in_ch, base, num_classes, num_pool, 2, 2, nn.Conv3d, norm_op, norm_kwargs, nn.Dropout3d,
                         {"p": 0, "inplace": True}, nn.LeakyReLU, {"negative_slope": 1e-2, "inplace": True}, True,
                         False, lambda x: x, G.InitWeights_He(1e-2), [[2, 2, 2]] * num_pool,
                         [[3, 3, 3]] * (num_pool + 1), False, True, True, max_num_features=max_num_features,
                         encoder_scale=encoder_scale)
    randomize_norm_params(net, seed + 1000)
    net.eval()
    net.do_ds = False
    if nonlin == "sigmoid":
        net.inference_apply_nonlin = nn.Sigmoid()  # BraTSRegions trainers
    elif nonlin == "softmax":
        net.inference_apply_nonlin = G.softmax_helper
    return net


def oracle_fns(net):
    """(forward_fn, nonlin) closures over the CPU oracle for `net`'s weights."""
    from oracle import unet as OU

    sd = {k: v.detach().cpu().float() for k, v in net.state_dict().items()}
    arch = OU.arch_from_module(net)
    return (lambda x: OU.forward(sd, arch, x)), sd, arch

"""Functional fp32 restatement of Generic_UNet.forward (reference model_architecture/generic_UNet.py:423-446).

Works from a plain ``state_dict`` (the reference module's key layout) plus a small architecture dict,
so it travels to the GPU box without /root/reference.  ``arch_from_module`` derives that dict from any module with
the reference layout (the reference class itself here, or the drop-in class of brainseg_b200).
"""
import torch
import torch.nn.functional as F


def arch_from_module(net):
    """Architecture facts the functional forward needs, read off the module tree (never hard-coded)."""
    import torch.nn as nn
    num_pool = len(net.tu)
    blk0 = net.conv_blocks_context[0].blocks[0]
    norm = blk0.instnorm
    if isinstance(norm, nn.BatchNorm3d):
        kind, groups = "bn", 0
    elif isinstance(norm, nn.GroupNorm):
        kind, groups = "gn", norm.num_groups
    elif isinstance(norm, nn.InstanceNorm3d):
        kind, groups = "in", 0
    else:
        raise ValueError(f"unsupported norm {type(norm)}")
    strides = []
    for d in range(num_pool + 1):
        stage = net.conv_blocks_context[d]
        first = stage.blocks[0] if d < num_pool else stage[0].blocks[0]
        strides.append(tuple(first.conv.stride))
    return {
        "num_pool": num_pool,
        "conv_per_stage": len(net.conv_blocks_context[0].blocks),
        "norm": kind,
        "groups": groups,
        "eps": float(norm.eps),
        "affine": bool(getattr(norm, "affine", True)),
        "slope": float(blk0.lrelu.negative_slope),
        "strides": strides,
        "num_classes": int(net.num_classes),
        "in_channels": int(blk0.conv.in_channels),
        "tu_strides": [tuple(t.stride) for t in net.tu],
    }


def _norm(x, sd, prefix, arch):
    w = sd.get(prefix + ".weight")
    b = sd.get(prefix + ".bias")
    if arch["norm"] == "bn":
        # eval-mode BatchNorm3d: running statistics (generic_UNet.py:65, network.eval() upstream)
        return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], w, b, False, 0.0,
                            arch["eps"])
    if arch["norm"] == "gn":
        return F.group_norm(x, arch["groups"], w, b, arch["eps"])
    return F.instance_norm(x, None, None, w, b, True, 0.0, arch["eps"])


def _block(x, sd, prefix, arch, stride):
    """ConvDropoutNormNonlin.forward, generic_UNet.py:68-72 (dropout is identity in eval)."""
    x = F.conv3d(x, sd[prefix + ".conv.weight"], sd.get(prefix + ".conv.bias"), stride=stride, padding=1)
    x = _norm(x, sd, prefix + ".instnorm", arch)
    return F.leaky_relu(x, arch["slope"])


def forward(sd, arch, x, return_features=False):
    """Generic_UNet.forward with do_ds=False (generic_UNet.py:423-446): returns final_nonlin-free logits of the
    last (full-resolution) head; the BraTS trainers build the net with final_nonlin = identity."""
    npool, cps = arch["num_pool"], arch["conv_per_stage"]
    skips = []
    with torch.no_grad():
        for d in range(npool):
            for i in range(cps):
                x = _block(x, sd, f"conv_blocks_context.{d}.blocks.{i}", arch, arch["strides"][d] if i == 0 else 1)
            skips.append(x)
        for i in range(cps - 1):
            x = _block(x, sd, f"conv_blocks_context.{npool}.0.blocks.{i}", arch,
                       arch["strides"][npool] if i == 0 else 1)
        x = _block(x, sd, f"conv_blocks_context.{npool}.1.blocks.0", arch, 1)
        for u in range(npool):
            x = F.conv_transpose3d(x, sd[f"tu.{u}.weight"], None, stride=arch["tu_strides"][u])
            x = torch.cat((x, skips[-(u + 1)]), dim=1)
            for i in range(cps - 1):
                x = _block(x, sd, f"conv_blocks_localization.{u}.0.blocks.{i}", arch, 1)
            x = _block(x, sd, f"conv_blocks_localization.{u}.1.blocks.0", arch, 1)
        feat = x
        logits = F.conv3d(x, sd[f"seg_outputs.{npool - 1}.weight"], sd.get(f"seg_outputs.{npool - 1}.bias"))
    if return_features:
        return logits, feat
    return logits


def conv_flops(sd, arch, patch):
    """Algorithmic 2*MAC count of one forward on a (D,H,W) patch, final head only."""
    npool, cps = arch["num_pool"], arch["conv_per_stage"]
    size = list(patch)
    total = 0.0

    def vox():
        return size[0] * size[1] * size[2]

    def conv(prefix, stride):
        nonlocal total
        w = sd[prefix + ".conv.weight"]
        for a in range(3):
            size[a] //= stride[a] if isinstance(stride, tuple) else stride
        total += 2.0 * w.shape[0] * w.shape[1] * 27 * vox()

    for d in range(npool):
        for i in range(cps):
            conv(f"conv_blocks_context.{d}.blocks.{i}", arch["strides"][d] if i == 0 else 1)
    for i in range(cps - 1):
        conv(f"conv_blocks_context.{npool}.0.blocks.{i}", arch["strides"][npool] if i == 0 else 1)
    conv(f"conv_blocks_context.{npool}.1.blocks.0", 1)
    for u in range(npool):
        w = sd[f"tu.{u}.weight"]
        total += 2.0 * w.shape[0] * w.shape[1] * 8 * vox()
        for a in range(3):
            size[a] *= arch["tu_strides"][u][a]
        for i in range(cps - 1):
            conv(f"conv_blocks_localization.{u}.0.blocks.{i}", 1)
        conv(f"conv_blocks_localization.{u}.1.blocks.0", 1)
    w = sd[f"seg_outputs.{npool - 1}.weight"]
    total += 2.0 * w.shape[0] * w.shape[1] * vox()
    return total

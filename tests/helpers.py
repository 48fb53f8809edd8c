"""Shared builders for the UNet parity tests."""
import torch
from torch import nn


from synthetic_case import build_dropin_unet  # noqa: E402,F401  (shared with bench.py)


def oracle_fns(net):
    """(forward_fn, nonlin) closures over the CPU oracle for `net`'s weights."""
    from oracle import unet as OU

    sd = {k: v.detach().cpu().float() for k, v in net.state_dict().items()}
    arch = OU.arch_from_module(net)
    return (lambda x: OU.forward(sd, arch, x)), sd, arch

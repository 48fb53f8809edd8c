"""Imports the REFERENCE's own modules from /root/reference (build container only) with stubs for the third-party
packages the image lacks (nnunet, axial_attention, nibabel).  Used by oracle/make_golden.py and by the
``needs_reference`` tests; never available on the GPU box, never imported by the product.

Stub provenance: the four nnunet/axial_attention symbols are import-time dependencies of
model_architecture/generic_UNet.py:17-24 that do not take part in Generic_UNet.forward for the BraTS trainers
(softmax_helper is replaced by the identity final_nonlin, InitWeights_He only draws the random initial weights,
SegmentationNetwork is the nn.Module base, the axial-attention classes are never instantiated when
axial_attention=False).  nibabel is only used for file I/O outside the functions we call.
"""
import importlib
import os
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("BSG_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model_architecture", "generic_UNet.py"))


class InitWeights_He(object):
    """nnU-Net v1 `InitWeights_He`: kaiming_normal_(a=neg_slope) on conv / transposed-conv weights, zero bias."""

    def __init__(self, neg_slope=1e-2):
        self.neg_slope = neg_slope

    def __call__(self, module):
        if isinstance(module, (nn.Conv3d, nn.Conv2d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
            module.weight = nn.init.kaiming_normal_(module.weight, a=self.neg_slope)
            if module.bias is not None:
                module.bias = nn.init.constant_(module.bias, 0)


def _install_stubs():
    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    for name in ("nnunet", "nnunet.utilities", "nnunet.network_architecture"):
        mod(name).__path__ = []
    mod("nnunet.utilities.nd_softmax").softmax_helper = lambda x: torch.softmax(x, 1)
    mod("nnunet.network_architecture.initialization").InitWeights_He = InitWeights_He

    class SegmentationNetwork(nn.Module):
        pass

    mod("nnunet.network_architecture.neural_network").SegmentationNetwork = SegmentationNetwork
    aa = mod("axial_attention")

    class _Absent(nn.Module):
        def __init__(self, *a, **k):
            raise RuntimeError("axial_attention is not part of the BraTS hot path")

    aa.AxialAttention = _Absent
    aa.AxialPositionalEmbedding = _Absent
    if "nibabel" not in sys.modules:
        nib = mod("nibabel")
        nib.load = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("nibabel stub: no file I/O in the oracle"))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


_cache = {}


def load_reference():
    """Returns a namespace with the reference modules on the hot path."""
    if "ns" in _cache:
        return _cache["ns"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import importlib.util  # noqa: F401

    _install_stubs()
    fe = os.path.join(REFERENCE_ROOT, "feature_extraction")
    if fe not in sys.path:
        sys.path.insert(0, fe)  # step3/step4 do `from utils import ...`
    ns = types.SimpleNamespace()
    ns.generic_UNet = _load(os.path.join(REFERENCE_ROOT, "model_architecture", "generic_UNet.py"), "ref_generic_UNet")
    ns.convert_labels = _load(os.path.join(REFERENCE_ROOT, "convert_labels_to_brats.py"), "ref_convert_labels")
    ns.evaluate = _load(os.path.join(REFERENCE_ROOT, "evaluate_segmentation.py"), "ref_evaluate_segmentation")
    ns.utils = importlib.import_module("utils")
    ns.step3 = _load(os.path.join(fe, "step3_multiplicity.py"), "ref_step3_multiplicity")
    ns.step4 = _load(os.path.join(fe, "step4_morphology.py"), "ref_step4_morphology")
    _cache["ns"] = ns
    return ns


def build_reference_unet(variant="bn", base=32, num_pool=5, in_ch=4, num_classes=3, seed=1, groups=8,
                         encoder_scale=1, max_num_features=None, randomize_norm=True):
    """Instantiates the REFERENCE Generic_UNet the way the BraTS-2021 V2 trainers do:
    Conv3d, dropout p=0, LeakyReLU(1e-2), deep supervision on, final_nonlin identity, conv pooling/upsampling."""
    ns = load_reference()
    G = ns.generic_UNet
    norm_op = {"bn": nn.BatchNorm3d, "in": nn.InstanceNorm3d, "gn": nn.GroupNorm}[variant]
    norm_kwargs = {"eps": 1e-5, "affine": True}
    if variant == "gn":
        norm_kwargs["num_groups"] = groups
    torch.manual_seed(seed)
    net = G.Generic_UNet(in_ch, base, num_classes, num_pool, 2, 2, nn.Conv3d, norm_op, norm_kwargs, nn.Dropout3d,
                         {"p": 0, "inplace": True}, nn.LeakyReLU, {"negative_slope": 1e-2, "inplace": True}, True,
                         False, lambda x: x, InitWeights_He(1e-2), [[2, 2, 2]] * num_pool,
                         [[3, 3, 3]] * (num_pool + 1), False, True, True, max_num_features=max_num_features,
                         encoder_scale=encoder_scale)
    if randomize_norm:
        randomize_norm_params(net, seed + 1000)
    net.eval()
    net.do_ds = False
    return net


from synthetic_case import randomize_norm_params  # noqa: E402,F401  (shared with bench.py / tests)

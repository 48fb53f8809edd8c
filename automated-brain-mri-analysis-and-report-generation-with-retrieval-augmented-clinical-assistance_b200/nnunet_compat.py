"""The slice of the nnU-Net v1 trainer API that the reference's inference script drives
(run_brats2021_inference_singlethread.py:21-23, 89-156, 178-183), implemented on the sm_100a engine:

    trainer, params = load_model_and_checkpoint_files(folder, folds, mixed_precision=True, checkpoint_name=...)
    d, s, dct = trainer.preprocess_patient(list_of_files)
    trainer.load_checkpoint_ram(params[k], False)
    seg, softmax = trainer.predict_preprocessed_data_return_seg_and_softmax(d, do_mirroring=..., mirror_axes=..., ...)
    save_segmentation_nifti_from_softmax(softmax, out_file, dct, order, region_class_order, ...)

The original classes live in the un-vendored `Brats21_KAIST_MRI_Lab/nnunet`.  The network
architecture is NOT hard-coded: it is inferred from the checkpoint's `state_dict` shapes (widths, depth, encoder scale,
norm type), so the KAIST variants (`..._BN_BD`, `..._largeUnet_Groupnorm`) load without their trainer source.
"""
import os
import pickle
import re

import numpy as np
import torch
from torch import nn

from . import nifti_io
from . import preprocessing
from .generic_UNet import Generic_UNet, InitWeights_He

DEFAULT_PATCH_SIZE = (128, 128, 128)  # Task500_BraTS2021 3d_fullres plans (data/temp_inference_output1)


def infer_network_config(state_dict, trainer_name="", num_groups=None):
    """Constructor arguments of Generic_UNet recovered from a checkpoint's state_dict."""
    sd = state_dict
    num_pool = len({int(m.group(1)) for k in sd for m in [re.match(r"tu\.(\d+)\.weight$", k)] if m})
    if num_pool == 0:
        raise ValueError("state_dict has no transposed-conv keys (tu.*.weight): not a Generic_UNet checkpoint")
    w_first = sd["conv_blocks_context.0.blocks.0.conv.weight"]
    in_ch, enc0 = int(w_first.shape[1]), int(w_first.shape[0])
    dec_last = int(sd[f"conv_blocks_localization.{num_pool - 1}.1.blocks.0.conv.weight"].shape[0])
    encoder_scale = max(1, enc0 // dec_last)
    base = enc0 // encoder_scale
    convs = len({int(m.group(1)) for k in sd for m in [re.match(r"conv_blocks_context\.0\.blocks\.(\d+)\.conv\.weight$", k)] if m})
    widths = [int(v.shape[0]) for k, v in sd.items() if k.endswith(".conv.weight")]
    num_classes = int(sd["seg_outputs.0.weight"].shape[0])
    prefix = "conv_blocks_context.0.blocks.0.instnorm."
    if prefix + "running_mean" in sd:
        norm_op, norm_kwargs = nn.BatchNorm3d, {"eps": 1e-5, "affine": True}
    elif prefix + "weight" in sd:
        if "groupnorm" in trainer_name.lower() or num_groups:
            norm_op, norm_kwargs = nn.GroupNorm, {"eps": 1e-5, "affine": True, "num_groups": int(num_groups or 8)}
        else:
            norm_op, norm_kwargs = nn.InstanceNorm3d, {"eps": 1e-5, "affine": True}
    else:
        norm_op, norm_kwargs = nn.InstanceNorm3d, {"eps": 1e-5, "affine": False}
    return {"input_channels": in_ch, "base_num_features": base, "num_classes": num_classes, "num_pool": num_pool,
            "num_conv_per_stage": convs, "norm_op": norm_op, "norm_op_kwargs": norm_kwargs,
            "max_num_features": max(widths), "encoder_scale": encoder_scale,
            "seg_output_use_bias": "seg_outputs.0.bias" in sd}


def build_network(cfg):
    """Generic_UNet the way the nnU-Net V2 trainers build it from an inferred configuration."""
    net = Generic_UNet(cfg["input_channels"], cfg["base_num_features"], cfg["num_classes"], cfg["num_pool"],
                       cfg["num_conv_per_stage"], 2, nn.Conv3d, cfg["norm_op"], cfg["norm_op_kwargs"], nn.Dropout3d,
                       {"p": 0, "inplace": True}, nn.LeakyReLU, {"negative_slope": 1e-2, "inplace": True}, True, False,
                       lambda x: x, InitWeights_He(1e-2), [[2, 2, 2]] * cfg["num_pool"],
                       [[3, 3, 3]] * (cfg["num_pool"] + 1), False, True, True,
                       max_num_features=cfg["max_num_features"], seg_output_use_bias=cfg["seg_output_use_bias"],
                       encoder_scale=cfg["encoder_scale"])
    net.eval()
    net.do_ds = False
    return net


class BratsTrainer:
    """Inference-only stand-in for nnUNetTrainerV2BraTSRegions_* (BraTSRegions trainers: sigmoid nonlinearity,
    regions_class_order (1, 2, 3))."""

    def __init__(self, plans, trainer_name="", regions=True, num_groups=None):
        self.plans = plans or {}
        self.trainer_name = trainer_name
        self.num_groups = num_groups
        self.network = None
        self.regions_class_order = (1, 2, 3) if regions else None
        self.data_aug_params = {"mirror_axes": (0, 1, 2)}
        self.patch_size = self._patch_size_from_plans()
        self.use_mask_for_norm = self._plan_value("use_mask_for_norm", {0: True})
        # this drop-in covers the BraTS geometry: identity axis order and images already at the plans' target spacing
        # (1 mm isotropic -> upstream's resampling is a no-op).  Anything else must fail loudly, not silently mis-scale.
        for key in ("transpose_forward", "transpose_backward"):
            axes = self._plan_value(key, [0, 1, 2])
            if list(axes) != [0, 1, 2]:
                raise NotImplementedError(f"plans[{key!r}] = {list(axes)}: only the identity axis order is supported")

    def _target_spacing(self):
        try:
            stages = self.plans["plans_per_stage"]
            return tuple(float(v) for v in stages[max(stages.keys())]["current_spacing"])
        except (KeyError, TypeError, ValueError):
            return None

    def _plan_value(self, key, default):
        return self.plans.get(key, default) if isinstance(self.plans, dict) else default

    def _patch_size_from_plans(self):
        try:
            stages = self.plans["plans_per_stage"]
            return tuple(int(v) for v in stages[max(stages.keys())]["patch_size"])
        except (KeyError, TypeError, ValueError):
            return DEFAULT_PATCH_SIZE

    # ---- checkpoints
    def load_checkpoint_ram(self, checkpoint, train=True):
        """checkpoint: the dict torch.load() returned for fold_k/model_final_checkpoint.model (key 'state_dict')."""
        sd = {k[7:] if k.startswith("module.") else k: v for k, v in checkpoint["state_dict"].items()}
        cfg = infer_network_config(sd, self.trainer_name, self.num_groups)
        key = {k: (v if not isinstance(v, dict) else tuple(sorted(v.items()))) for k, v in cfg.items()}
        if self.network is None or getattr(self, "_cfg_key", None) != key:
            self.network = build_network(cfg)
            self._cfg_key = key
        self.network.load_state_dict(sd)  # re-packs the weights into the cached engines in place
        self.network.eval()
        self.network.inference_apply_nonlin = nn.Sigmoid() if self.regions_class_order is not None else (
            lambda x: torch.softmax(x, 1))

    # ---- preprocessing
    def preprocess_patient(self, input_files):
        """Reads the modalities (z, y, x), crops to the non-zero box and z-scores each channel inside the mask on the
        device.  Returns (d, s, properties) like upstream: `d` a cuda fp32 tensor (C, z, y, x), `s` None."""
        imgs = [nifti_io.load(f) for f in input_files]
        data = np.stack([im.get_fdata().astype(np.float32) for im in imgs])
        if isinstance(self.use_mask_for_norm, dict):  # per channel, as upstream's GenericPreprocessor applies it
            use_mask = [bool(self.use_mask_for_norm.get(c, self.use_mask_for_norm.get(str(c), True)))
                        for c in range(data.shape[0])]
        else:
            use_mask = bool(self.use_mask_for_norm)
        target = self._target_spacing()
        if target is not None:
            have = tuple(float(v) for v in imgs[0].zooms[:3])[::-1]  # NIfTI (x, y, z) zooms -> array order (z, y, x)
            if not np.allclose(have, target, atol=1e-3):
                raise NotImplementedError(f"image spacing {have} differs from the plans' target spacing {target}: "
                                          "resampling is outside this drop-in (BraTS data is 1 mm isotropic)")
        d, props = preprocessing.preprocess_case(data, use_mask_for_norm=use_mask)
        props["list_of_data_files"] = list(input_files)
        props["nifti_like"] = imgs[0]
        props["itk_spacing"] = imgs[0].zooms
        return d, None, props

    # ---- prediction
    def predict_preprocessed_data_return_seg_and_softmax(self, data, do_mirroring=True, mirror_axes=None,
                                                         use_sliding_window=True, step_size=0.5, use_gaussian=True,
                                                         pad_border_mode="constant", pad_kwargs=None, all_in_gpu=False,
                                                         verbose=True, mixed_precision=True):
        if pad_border_mode == "constant" and pad_kwargs is None:
            pad_kwargs = {"constant_values": 0}
        if mirror_axes is None:
            mirror_axes = self.data_aug_params["mirror_axes"]
        return self.network.predict_3D(data, do_mirroring, mirror_axes, use_sliding_window, step_size, self.patch_size,
                                       self.regions_class_order, use_gaussian, pad_border_mode, pad_kwargs, all_in_gpu,
                                       verbose, mixed_precision)


def load_model_and_checkpoint_files(folder, folds=None, mixed_precision=None, checkpoint_name="model_best"):
    """`<folder>/plans.pkl` + `<folder>/fold_k/<checkpoint_name>.model` (run_brats2021_inference_singlethread.py:253-264).
    Returns (trainer, [checkpoint dicts])."""
    if folds is None:
        folds = sorted(int(d[5:]) for d in os.listdir(folder) if d.startswith("fold_") and d[5:].isdigit())
    elif isinstance(folds, int):
        folds = [folds]
    plans = {}
    plans_file = os.path.join(folder, "plans.pkl")
    if os.path.isfile(plans_file):
        with open(plans_file, "rb") as f:
            plans = pickle.load(f)
    trainer_name = os.path.basename(os.path.normpath(str(folder))).split("__")[0]
    trainer = BratsTrainer(plans, trainer_name, regions="Regions" in trainer_name or not trainer_name)
    params = []
    for k in folds:
        path = os.path.join(folder, f"fold_{k}", f"{checkpoint_name}.model")
        if not os.path.isfile(path):
            raise FileNotFoundError(f"checkpoint missing: {path}")
        params.append(torch.load(path, map_location="cpu", weights_only=False))
    return trainer, params


def save_segmentation_nifti_from_softmax(segmentation_softmax, out_fname, properties_dict, order=1,
                                         region_class_order=None, seg_postprogess_fn=None, seg_postprocess_args=None,
                                         resampled_npz_fname=None, non_postprocessed_fname=None, force_separate_z=None,
                                         interpolation_order_z=0, verbose=True):
    """decide labels (argmax, or the ordered `> 0.5` assignment for regions), paste into the original
    geometry at `crop_bbox`, write uint8 NIfTI with the source image's header.  1 mm BraTS data: shapes after cropping
    and prediction agree, so no resampling happens (a mismatch raises)."""
    probs = segmentation_softmax
    if torch.is_tensor(probs):
        probs = probs.cpu().numpy()
    if tuple(probs.shape[1:]) != tuple(properties_dict["size_after_cropping"]):
        raise NotImplementedError("resampling of the softmax is not needed for 1 mm isotropic BraTS data")
    if region_class_order is None:
        seg = probs.argmax(0)
    else:
        seg = np.zeros(probs.shape[1:], dtype=np.float32)
        for i, c in enumerate(region_class_order):
            seg[probs[i] > 0.5] = c
    full = preprocessing.uncrop_segmentation(torch.from_numpy(np.ascontiguousarray(seg.astype(np.uint8))),
                                             properties_dict).numpy()
    nifti_io.save(out_fname, full, properties_dict["nifti_like"])
    return out_fname

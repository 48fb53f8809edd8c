"""Drop-in for nnU-Net v1 `SegmentationNetwork` (UPSTREAM nnunet/network_architecture/neural_network.py; the base
class of the reference's Generic_UNet, model_architecture/generic_UNet.py:22,171): `predict_3D` with the upstream
signature and return convention, executed by the sm_100a engine.
"""
import numpy as np
import torch
from torch import nn

from . import sliding


class SegmentationNetwork(nn.Module):
    def __init__(self):
        super().__init__()
        self.input_shape_must_be_divisible_by = None
        self.conv_op = None
        self.num_classes = None
        self.inference_apply_nonlin = lambda x: x  # upstream default; trainers install sigmoid / softmax_helper
        self._engines = {}
        self.engine_batch = 8  # (tile, mirror) forwards in flight, split evenly over `engine_lanes` CUDA streams
        self.engine_lanes = 2  # >1: HBM-bound passes of one lane overlap the tensor-bound convs of the other
        # None: BSG_ACT_DTYPE / auto (fp16); "bf16" after the fp16 range guard fired; "fp32": fp32-equivalent arithmetic
        # (fp16x3 split operands, engine.py) — what predict_3D(mixed_precision=False) selects, like upstream's no-autocast path
        self.engine_dtype = None

    # ------------------------------------------------------------------ engine cache
    def engine_for(self, patch_size, batch=None, slot=0, dtype=None):
        """Cached UNetEngine for a patch size / batch; `slot` distinguishes the engines of concurrent stream lanes;
        `dtype` overrides the network's engine_dtype for this engine."""
        from . import _lib as L
        from .engine import UNetEngine
        if not torch.cuda.is_available():
            raise L.BsgError("brainseg_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        batch = int(batch or self.engine_batch)
        dtype = dtype or self.engine_dtype
        key = (tuple(int(p) for p in patch_size), batch, torch.cuda.current_device(), int(slot), dtype)
        if key not in self._engines:
            self._engines[key] = UNetEngine(self, key[0], batch, act_dtype=dtype)
        return self._engines[key]

    def engines_for(self, patch_size, batch=None, lanes=None, dtype=None):
        """The lane engines of the sliding-window driver: `lanes` engines of batch/lanes forwards each."""
        batch = int(batch or self.engine_batch)
        lanes = max(1, min(int(lanes or self.engine_lanes), batch))
        per = -(-batch // lanes)
        return [self.engine_for(patch_size, per, slot, dtype) for slot in range(lanes)]

    def invalidate_engines(self):
        """Drops every cached engine (device buffers, plans, tensor maps) — after a change of ARCHITECTURE or of the
        activation dtype.  New weights alone do not need it: see refresh_engines()."""
        for eng in self._engines.values():
            eng.close()
        self._engines = {}

    def refresh_engines(self):
        """Call after the parameters changed in place (load_state_dict / load_checkpoint_ram per fold): re-packs the
        weights into the cached engines' existing device tensors; activation buffers, plans and tensor maps stay."""
        for eng in self._engines.values():
            eng.reload_weights()

    def use_bf16_activations(self):
        """Answer to the fp16 range guard: re-plan this network with bf16 activations and weights (8 exponent bits)."""
        if self.engine_dtype != "bf16":
            self.engine_dtype = "bf16"
            self.invalidate_engines()

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.refresh_engines()
        return out

    def _nonlin_name(self):
        f = self.inference_apply_nonlin
        if isinstance(f, nn.Sigmoid):
            return "sigmoid"
        if isinstance(f, nn.Softmax) or getattr(f, "__name__", "") == "softmax_helper":
            return "softmax"
        probe = torch.tensor([[0.5, -1.0]])
        if torch.equal(f(probe), probe):
            return "identity"
        if torch.allclose(f(probe), torch.sigmoid(probe)):
            return "sigmoid"
        if torch.allclose(f(probe), torch.softmax(probe, 1)):
            return "softmax"
        raise NotImplementedError("inference_apply_nonlin must be sigmoid, softmax over dim 1 or identity")

    # ------------------------------------------------------------------ upstream API
    def predict_3D(self, x, do_mirroring, mirror_axes=(0, 1, 2), use_sliding_window=False, step_size=0.5,
                   patch_size=None, regions_class_order=None, use_gaussian=False, pad_border_mode="constant",
                   pad_kwargs=None, all_in_gpu=False, verbose=True, mixed_precision=True):
        """x: (c, z, y, x) float array.  Returns (seg, class_probabilities) as numpy arrays like upstream:
        seg (z, y, x) — argmax int64, or float32 region labels when `regions_class_order` is given;
        class_probabilities float32 (num_classes, z, y, x)."""
        assert step_size <= 1, "step_size must be smaller than 1. Otherwise there will be a gap between consecutive predictions"
        if pad_border_mode != "constant" or (pad_kwargs not in (None, {"constant_values": 0})):
            raise NotImplementedError("only constant zero padding (the nnU-Net default used by the reference)")
        if len(mirror_axes) and max(mirror_axes) > 2:
            raise ValueError("mirror axes. duh")
        x = np.asarray(x) if not torch.is_tensor(x) else x
        assert len(x.shape) == 4, "data must have shape (c,x,y,z)"
        # upstream: mixed_precision=True runs the forwards under CUDA autocast (fp16), False in fp32
        dtype = None if mixed_precision else "fp32"
        if use_sliding_window:
            assert patch_size is not None, "patch_size cannot be None for tiled prediction"
            seg, probs = self.predict_3D_device(x, do_mirroring, mirror_axes, step_size, patch_size,
                                                regions_class_order, use_gaussian, dtype=dtype)
        else:
            # upstream _internal_predict_3D_3Dconv: pad to >= patch_size (its `min_size`) and to a multiple of
            # input_shape_must_be_divisible_by, ONE mirrored forward over the whole volume (no Gaussian), crop back —
            # i.e. a single tile that is the padded volume itself
            div = [int(d) for d in self.input_shape_must_be_divisible_by]
            floor = [int(p) for p in patch_size] if patch_size is not None else [0, 0, 0]
            whole = [max(s, f) for s, f in zip(x.shape[1:], floor)]
            whole = [w if w % d == 0 else w + d - w % d for w, d in zip(whole, div)]
            seg, probs = self.predict_3D_device(x, do_mirroring, mirror_axes, 1.0, whole, regions_class_order, False,
                                                engine_batch=1, dtype=dtype)
        seg = seg.cpu().numpy()
        seg = seg.astype(np.float32) if regions_class_order is not None else seg.astype(np.int64)
        return seg, probs.cpu().numpy()

    def predict_3D_device(self, x, do_mirroring=True, mirror_axes=(0, 1, 2), step_size=0.5, patch_size=None,
                          regions_class_order=None, use_gaussian=True, want_probs=True, engine_batch=None, dtype=None):
        """predict_3D without the final device->host copies: returns (uint8 seg, fp32 probs) cuda tensors."""
        dev = torch.device("cuda", torch.cuda.current_device())
        vol = (torch.from_numpy(np.ascontiguousarray(x)) if not torch.is_tensor(x) else x).to(dev, torch.float32)
        patch = tuple(int(p) for p in patch_size)
        # pad_nd_image(x, patch, 'constant', 0): symmetric zero pad up to the patch size, floor on the low side
        shape = tuple(vol.shape[1:])
        new = [max(s, p) for s, p in zip(shape, patch)]
        pads, lo = [], []
        for s, n in zip(shape, new):
            lo.append((n - s) // 2)
            pads.append(((n - s) // 2, (n - s) // 2 + (n - s) % 2))
        if any(a or b for a, b in pads):
            vol = torch.nn.functional.pad(vol, (pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]))
        vol = vol.contiguous()
        codes = sliding.mirror_codes_for(mirror_axes, do_mirroring)
        pred = sliding.SlidingWindowPredictor(self.engines_for(patch, engine_batch, dtype=dtype), step_size, use_gaussian, codes,
                                              self._nonlin_name())
        acc = pred.accumulate(vol)
        seg, probs = pred.finalize([acc], tuple(vol.shape[1:]), regions_class_order, want_probs)
        if tuple(vol.shape[1:]) != shape:  # crop back with the slicer pad_nd_image returned
            sl = tuple(slice(l, l + s) for l, s in zip(lo, shape))
            seg = seg[sl].contiguous()
            if probs is not None:
                probs = probs[(slice(None),) + sl].contiguous()
        self.last_kernel_launches = pred.kernel_launches
        return seg, probs

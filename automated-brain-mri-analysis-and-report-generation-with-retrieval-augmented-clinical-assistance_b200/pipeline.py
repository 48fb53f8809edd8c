"""In-memory case pipeline: the hot path of run_brats2021_inference_singlethread.main (:246-322) and the voxel steps
of run_full_pipeline.py that consume its output (convert_labels :198, evaluate :226, feature extraction :274), with
the label volume staying on the device between stages.

    model 1 -> probs -> regions (1,2,3) -> seg1        predict_case_single_threaded :81-158
    model 2 -> probs -> regions (1,2,3) -> seg2
    ensemble = np.round((seg1 + seg2) / 2)              main :305
    brats    = convert_labels_to_brats2025(ensemble)    convert_labels_to_brats.py:34-43 (pipeline default format)
    Dice vs GT, connected components, morphology        evaluate_segmentation.py, feature_extraction/step3, step4
"""
import numpy as np
import torch

from . import sliding
from . import voxelops as V
from . import convert_labels_to_brats as CL
from . import evaluate_segmentation as EV
from .feature_extraction import step3_multiplicity as S3
from .feature_extraction import step4_morphology as S4
from .feature_extraction import utils as FU


class BratsCasePipeline:
    def __init__(self, models, patch_size=(128, 128, 128), step_size=0.5, mirror_axes=(0, 1, 2), do_mirroring=True,
                 use_gaussian=True, regions_class_order=(1, 2, 3), label_format="brats2025", batch=8, rank=0,
                 world_size=1, reduce_fn=None, lanes=None, ensemble="label_round", small_et_threshold=None,
                 small_et_replace=2):
        """models: one entry per ensemble member — a drop-in Generic_UNet, or a list of them = the folds of that model,
        whose probabilities are averaged before the decision (np.mean over folds, reference :128);
        `reduce_fn(acc)` sums an accumulator over ranks when the (tile, mirror) work items of ONE case are sharded
        (latency mode)."""
        if ensemble not in ("label_round", "prob_mean"):
            raise ValueError("ensemble must be 'label_round' (run_brats2021_inference_singlethread.py:305) or 'prob_mean' "
                             "(archived/kaist_original_inference.py:29-33)")
        # 'prob_mean' = the original KAIST pipeline: nnUNet_ensemble (mean of the members' probabilities, then the regions
        # decision) + apply_threshold_to_folder(..., small_et_threshold=200, small_et_replace=2): an enhancing-tumour label
        # (3, nnU-Net convention) with fewer voxels than the threshold is relabelled
        self.ensemble, self.small_et_threshold, self.small_et_replace = ensemble, small_et_threshold, small_et_replace
        self.models = [list(m) if isinstance(m, (list, tuple)) else [m] for m in models]
        self.patch = tuple(patch_size)
        self.regions = regions_class_order
        self.lut = CL.LUT_BRATS2025 if label_format == "brats2025" else CL.LUT_BRATS2021
        self.reduce_fn = reduce_fn
        codes = sliding.mirror_codes_for(mirror_axes, do_mirroring)
        self.fold_predictors = []  # [model][fold]
        for folds in self.models:
            row = []
            for net in folds:
                engs = net.engines_for(self.patch, batch, lanes)
                row.append(sliding.SlidingWindowPredictor(engs, step_size, use_gaussian, codes, net._nonlin_name(), rank,
                                                          world_size))
            self.fold_predictors.append(row)
        self.predictors = [row[0] for row in self.fold_predictors]  # first fold of every model (bench / diagnostics)
        self.device = self.predictors[0].device
        self._post_stream = None

    def kernel_launches(self):
        return sum(p.kernel_launches for row in self.fold_predictors for p in row)

    def segment(self, vol):
        """vol: fp32 cuda tensor (C, Z, Y, X), extents >= patch.  Returns the per-model uint8 label volumes."""
        shape = tuple(vol.shape[1:])
        segs = []
        if self.ensemble == "prob_mean":
            # every member (model x fold) weighs the same, as np.mean over the folds' means does for equal fold counts
            accs = []
            for row in self.fold_predictors:
                for pred in row:
                    acc = pred.accumulate(vol)
                    accs.append(self.reduce_fn(acc) if self.reduce_fn is not None else acc)
            seg, _ = self.fold_predictors[0][0].finalize(accs, shape, self.regions, want_probs=False)
            return [seg]
        for row in self.fold_predictors:
            accs = []
            for pred in row:
                acc = pred.accumulate(vol)
                if self.reduce_fn is not None:
                    acc = self.reduce_fn(acc)
                accs.append(acc)
            seg, _ = row[0].finalize(accs, shape, self.regions, want_probs=False)  # mean over folds, then decide
            segs.append(seg)
        return segs

    # ------------------------------------------------------------------ two-phase API (cohort throughput)
    # submit() enqueues the inference of a case and returns at once; finish() runs the post-processing on a second
    # stream, whose host-side glue and small device->host reads then overlap the inference of the NEXT submitted case:
    #     h = pipe.submit(case[0]); for i: nxt = pipe.submit(case[i+1]); out[i] = pipe.finish(h); h = nxt
    def submit(self, volume, gt=None):
        """volume: (C, Z, Y, X) float32, host (numpy / pinned tensor) or device; gt: optional label volume.  Enqueues
        host->device copies, both models' sliding-window inference and the ensemble + remap; no host synchronisation."""
        if isinstance(volume, np.ndarray):
            volume = torch.from_numpy(volume)
        vol = volume.to(self.device, torch.float32, non_blocking=True).contiguous()
        segs = self.segment(vol)
        if len(segs) == 1:
            seg = segs[0]
            if self.small_et_threshold is not None:
                # apply_brats_threshold: `if np.sum(seg == 3) < threshold: seg[seg == 3] = replace_with`; the count is
                # needed on the host to pick the LUT (one 120-byte read)
                n_et = int(V.masked_moments(seg, [V.bits_of(3)])[0]["count"])
                if n_et < self.small_et_threshold:
                    lut = np.array(self.lut, dtype=np.uint8).copy()
                    lut[3] = self.lut[self.small_et_replace]
                    brats = V.label_lut(seg, lut)
                else:
                    brats = V.label_lut(seg, self.lut)
            else:
                brats = V.label_lut(seg, self.lut)
        elif len(segs) == 2:
            brats = V.ensemble_round(segs[0], segs[1], post_lut=self.lut)  # ensemble + remap in one pass
        else:
            raise NotImplementedError("the reference ensembles exactly two models")
        gt_dev = None
        if gt is not None:
            gt_dev = V.as_label_volume(gt)
            if tuple(gt_dev.shape) != tuple(brats.shape):
                gt_dev = gt  # let evaluate_arrays report the mismatch the way the reference does
        done = torch.cuda.Event()
        done.record()
        return {"vol": vol, "segs": segs, "brats": brats, "gt": gt_dev, "done": done}

    def finish(self, pending, voxel_dims=(1.0, 1.0, 1.0), features=True):
        """Post-processing of a submitted case (Dice, components, morphology) on the pipeline's second stream."""
        if self._post_stream is None:
            self._post_stream = torch.cuda.Stream(self.device)
        s = self._post_stream
        s.wait_event(pending["done"])
        brats, segs, gt = pending["brats"], pending["segs"], pending["gt"]
        for t in [brats] + list(segs) + ([gt] if torch.is_tensor(gt) and gt.is_cuda else []):
            t.record_stream(s)
        out = {"segmentation": brats, "model_segmentations": segs}
        self.extra_launches = 1
        with torch.cuda.stream(s):
            if gt is not None:
                out["evaluation"] = EV.evaluate_arrays(brats, gt)
                self.extra_launches += 1
            if features:
                lv = FU.LabelVolume(brats)
                masks = FU.get_tumor_masks(lv)
                out["components"] = S3.detect_connected_components(lv, voxel_dims)
                out["enhancing"] = S3.analyze_enhancing_components(lv, voxel_dims)
                out["shape"] = S4.calculate_shape_descriptors(lv, masks, voxel_dims)
                out["necrosis"] = S4.analyze_necrosis_pattern(lv, masks, np.array(voxel_dims))
                vv = float(np.prod(voxel_dims)) / 1000.0
                out["volumes_cm3"] = {k: FU.calculate_volume(masks[k], vv) for k in ("ncr", "ed", "et", "tc", "wt")}
                self.extra_launches += 2 * 8 + 2
            s.synchronize()
        return out

    def run_case(self, volume, gt=None, voxel_dims=(1.0, 1.0, 1.0), features=True):
        """One case start to finish: submit() + finish().  Returns a dict with the ensemble label volume in BraTS
        convention (device uint8) and the scalar results of the post-processing steps."""
        return self.finish(self.submit(volume, gt), voxel_dims, features)

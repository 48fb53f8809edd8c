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
                 small_et_replace=2, shard=None):
        """models: one entry per ensemble member — a drop-in Generic_UNet, or a list of them = the folds of that model,
        whose probabilities are averaged before the decision (np.mean over folds, reference :128);
        `reduce_fn(acc)` sums an accumulator over ranks when the (tile, mirror) work items of ONE case are sharded
        (latency mode); `shard` (a sharded.ShardedExchange) replaces it by the peer-memory / NCCL exchange of that
        module: accumulators and label volumes then live in buffers the exchange owns."""
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
        self.shard = shard
        self._build_args = (step_size, use_gaussian, batch, rank, world_size, lanes)
        codes = sliding.mirror_codes_for(mirror_axes, do_mirroring)
        self._codes = codes
        self._build_predictors()
        self._post_stream = None
        self._copy_stream = None
        self.extra_launches = 0
        self.fp16_overflows = 0  # times the fp16 range guard switched the networks to bf16 engines
        self._gen = 0            # bumped at every such switch: cases submitted before it are re-run

    def _build_predictors(self):
        step_size, use_gaussian, batch, rank, world_size, lanes = self._build_args
        codes = self._codes
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

    def kernel_launches(self):
        return sum(p.kernel_launches for row in self.fold_predictors for p in row)

    def segment(self, vol):
        """vol: fp32 cuda tensor (C, Z, Y, X), extents >= patch.  Returns the per-model uint8 label volumes."""
        shape = tuple(vol.shape[1:])
        segs = []
        if self.shard is not None:
            return self._segment_sharded(vol, shape)
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

    def _segment_sharded(self, vol, shape):
        """One case over all ranks: this rank's share of every member's (tile, mirror) forwards goes into accumulators
        the exchange owns; each ensemble member's exchange + finalize runs on a side stream while the next member's
        forwards are already being issued; every rank ends up with the complete label volumes."""
        sh = self.shard
        main = torch.cuda.current_stream(self.device)
        side = sh.side_stream()
        trace = getattr(self, "trace", None)  # bench: list collecting (label, event) pairs of one case

        def mark(label, stream):
            if trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                trace.append((label, ev))

        mark("start", main)
        members = [[p for row in self.fold_predictors for p in row]] if self.ensemble == "prob_mean" else \
            [list(row) for row in self.fold_predictors]
        segs = []
        for mi, row in enumerate(members):
            keys = []
            for fi, pred in enumerate(row):
                key = ("acc", mi, fi, shape)
                acc, _ = sh.shared(key, (pred.engine.num_classes,) + shape, torch.float32)
                acc.zero_()
                pred.accumulate(vol, acc)
                keys.append(key)
            mark(f"forwards_{mi}", main)
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
            _, _, wsum = row[0].geometry(shape)
            with torch.cuda.stream(side):
                seg = sh.finalize(keys, wsum, row[0].engine.num_classes, self.regions, ("seg", mi, shape))
            mark(f"exchange_{mi}", side)
            segs.append(seg.view(shape))
        done = torch.cuda.Event()
        done.record(side)
        main.wait_event(done)
        return segs

    def _engines(self):
        return [eng for row in self.fold_predictors for pred in row for eng in pred.engines]

    # ------------------------------------------------------------------ two-phase API (cohort throughput)
    # submit() enqueues the inference of a case and returns at once; finish() runs the post-processing on a second
    # stream, whose host-side glue and small device->host reads then overlap the inference of the NEXT submitted case:
    #     h = pipe.submit(case[0]); for i: nxt = pipe.submit(case[i+1]); out[i] = pipe.finish(h); h = nxt
    def submit(self, volume, gt=None):
        """volume: (C, Z, Y, X) float32, host (numpy / pinned tensor) or device; gt: optional label volume.  Enqueues
        host->device copies, both models' sliding-window inference and the ensemble + remap; no host synchronisation."""
        if isinstance(volume, np.ndarray):
            volume = torch.from_numpy(volume)
        if volume.is_cuda:
            vol = volume.to(self.device, torch.float32).contiguous()
        else:
            # host -> device on a copy stream: the upload of case i+1 overlaps the kernels of case i still in flight
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(self.device)
            main = torch.cuda.current_stream(self.device)
            with torch.cuda.stream(self._copy_stream):
                vol = volume.to(self.device, torch.float32, non_blocking=True).contiguous()
                up = torch.cuda.Event()
                up.record(self._copy_stream)
            main.wait_event(up)
            vol.record_stream(main)
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
        elif len(segs) != 2:
            raise NotImplementedError("the reference ensembles exactly two models")
        gt_dev, hist_buf = None, None
        if gt is not None:
            gt_dev = V.as_label_volume(gt)
            if tuple(gt_dev.shape) != tuple(segs[0].shape):
                gt_dev = gt  # let evaluate_arrays report the mismatch the way the reference does
        if len(segs) == 2:
            if torch.is_tensor(gt_dev) and gt_dev.is_cuda and gt_dev.dtype == torch.uint8 and tuple(gt_dev.shape) == tuple(segs[0].shape):
                # ensemble + remap + Dice bins against the ground truth in one pass over the label volumes
                brats, hist_buf = V.ensemble_remap_hist(segs[0], segs[1], gt_dev, post_lut=self.lut)
            else:
                brats = V.ensemble_round(segs[0], segs[1], post_lut=self.lut)  # ensemble + remap in one pass
        done = torch.cuda.Event()
        done.record()
        return {"vol": vol, "segs": segs, "brats": brats, "gt": gt_dev, "gt_in": gt, "done": done, "gen": self._gen,
                "hist": hist_buf}

    def finish(self, pending, voxel_dims=(1.0, 1.0, 1.0), features=True, post=True, labels=None):
        """Post-processing of a submitted case (Dice, components, morphology) on the pipeline's second stream.
        `post=False` (sharded mode: a rank that does not own this case's post-processing) only waits for the case.
        `labels` (device uint8 volume): run the post-processing on THIS label volume instead of the case's own
        output — benchmarks of the post-processing leg on realistic (blobby) label volumes; the returned
        `segmentation` is still the case's own."""
        if self._post_stream is None:
            self._post_stream = torch.cuda.Stream(self.device)
        s = self._post_stream
        s.wait_event(pending["done"])
        if self._guarded() or pending["gen"] != self._gen:
            # fp16 range guard: a conv epilogue stored a value beyond 65504 -> label volumes computed by the fp16 engines
            # are not trustworthy.  Re-plan the networks in bf16 (8 exponent bits) and run the case again — this one and
            # any case submitted before the switch (their `gen` is stale).
            from .engine import overflow_flag
            hit = False
            if self._guarded():
                with torch.cuda.stream(s):
                    hit = bool(overflow_flag(self.device).item())  # 4-byte read on the post stream, after `done`
            if hit:
                torch.cuda.synchronize(self.device)  # nothing in flight may still use the engines being dropped
                overflow_flag(self.device).zero_()
                self.fp16_overflows += 1
                self._gen += 1
                for folds in self.models:
                    for net in folds:
                        net.use_bf16_activations()
                self._build_predictors()
            if hit or pending["gen"] != self._gen:
                pending = self.submit(pending["vol"], pending["gt_in"])
                s.wait_event(pending["done"])
        if not post:
            pending["done"].synchronize()
            return {"segmentation": pending["brats"], "model_segmentations": pending["segs"]}
        brats, segs, gt = pending["brats"], pending["segs"], pending["gt"]
        for t in [brats] + list(segs) + ([gt] if torch.is_tensor(gt) and gt.is_cuda else []):
            t.record_stream(s)
        out = {"segmentation": brats, "model_segmentations": segs}
        if labels is not None:
            if tuple(labels.shape) != tuple(brats.shape):
                labels = labels.reshape(brats.shape) if labels.numel() == brats.numel() else labels
            brats = labels
        self.extra_launches = 1
        with torch.cuda.stream(s):
            if gt is not None:
                hist = None
                if pending.get("hist") is not None and labels is None:
                    pending["hist"].record_stream(s)
                    hist = V.joint_hist_from_buffer(pending["hist"])
                out["evaluation"] = EV.evaluate_arrays(brats, gt, _hist=hist)
                self.extra_launches += 0 if hist is not None else 1
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

    def _guarded(self):
        return any(eng.guard for eng in self._engines())

    def run_case(self, volume, gt=None, voxel_dims=(1.0, 1.0, 1.0), features=True):
        """One case start to finish: submit() + finish().  Returns a dict with the ensemble label volume in BraTS
        convention (device uint8) and the scalar results of the post-processing steps."""
        return self.finish(self.submit(volume, gt), voxel_dims, features)

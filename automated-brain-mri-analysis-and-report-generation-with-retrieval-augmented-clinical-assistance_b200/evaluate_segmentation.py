"""Drop-in for the metric functions of the reference's evaluate_segmentation.py.

One joint-histogram kernel (bsg_joint_hist_u8) replaces the ~20 full-volume float32 temporaries per label; the
scalar formulas (:35-38, :187-189) are evaluated on np.float32 counts exactly as the reference does, so every
returned value is bit-identical (counts stay below 2**24, where float32 sums are exact).
"""
import numpy as np

from . import voxelops as V

WT_LABELS, TC_LABELS = (1, 2, 3), (1, 3)  # evaluate_segmentation.py:130-131, :140-141


class LabelPairHistogram:
    """Joint histogram of (prediction, ground truth), computed once on the device."""

    def __init__(self, pred, gt):
        p, g = V.as_label_volume(pred), V.as_label_volume(gt)
        if p.shape != g.shape:
            raise ValueError("shape mismatch")
        self.hist = V.joint_hist(p, g)
        self.total = int(self.hist.sum())

    def counts(self, labels):
        """(tp, fp, fn, tn) as np.float32 for the binary masks isin(pred, labels) / isin(gt, labels)."""
        sel = np.zeros(16, dtype=bool)
        for l in labels:
            if float(l) == int(l) and 0 <= int(l) < 16:
                sel[int(l)] = True
        h = self.hist
        tp = int(h[np.ix_(sel, sel)].sum())
        fp = int(h[sel, :].sum()) - tp
        fn = int(h[:, sel].sum()) - tp
        tn = self.total - tp - fp - fn
        return np.float32(tp), np.float32(fp), np.float32(fn), np.float32(tn)

    def present_labels(self):
        """sorted(set(np.unique(pred)) | set(np.unique(gt))) (:84-105)."""
        present = (self.hist.sum(axis=0) + self.hist.sum(axis=1)) > 0
        return [l for l in range(16) if present[l]]


def _metrics(tp, fp, fn, tn):
    dice = (2 * tp) / (2 * tp + fp + fn + 1e-8)
    iou = tp / (tp + fp + fn + 1e-8)
    sensitivity = tp / (tp + fn + 1e-8)
    specificity = tn / (tn + fp + 1e-8)
    return {"dice": dice, "iou": iou, "sensitivity": sensitivity, "specificity": specificity,
            "tp": tp, "fp": fp, "fn": fn, "tn": tn}


def calculate_metrics(pred, gt, label, _hist=None):
    """Per-label Dice / IoU / sensitivity / specificity and TP/FP/FN/TN (reference :12-49)."""
    h = _hist or LabelPairHistogram(pred, gt)
    return _metrics(*h.counts([label]))


def calculate_metrics_binary(pred_mask, gt_mask, _counts=None):
    """Metrics for two binary masks (reference :181-195)."""
    if _counts is None:
        _counts = LabelPairHistogram(pred_mask, gt_mask).counts([1])
    m = _metrics(*_counts)
    return {"dice": m["dice"], "iou": m["iou"], "sensitivity": m["sensitivity"]}


def evaluate_arrays(pred_data, gt_data):
    """The arithmetic of evaluate_segmentation() (:84-162) on in-memory label volumes.

    Returns None on a shape mismatch like the reference (:78-81), else
    {"labels": {label: metrics}, "wt": ..., "tc": ..., "et": metrics | None, "mean_dice": ...}."""
    if tuple(pred_data.shape) != tuple(gt_data.shape):
        print("\n⚠️  WARNING: Shape mismatch! Attempting to resize...")
        return None
    h = LabelPairHistogram(pred_data, gt_data)
    all_metrics = {}
    for label in h.present_labels():
        if label == 0:
            continue
        all_metrics[label] = calculate_metrics(None, None, label, _hist=h)
    wt = calculate_metrics_binary(None, None, _counts=h.counts(WT_LABELS))
    tc = calculate_metrics_binary(None, None, _counts=h.counts(TC_LABELS))
    et = all_metrics.get(3)
    mean_dice = np.mean([wt["dice"], tc["dice"], et["dice"] if et is not None else 0])
    return {"labels": all_metrics, "wt": wt, "tc": tc, "et": et, "mean_dice": mean_dice}

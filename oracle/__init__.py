"""CPU oracle for the brainseg_b200 hot path — TEST INFRASTRUCTURE ONLY.

A plain torch-fp32 / numpy / scipy restatement of the reference's algorithm for the hot path (Generic_UNet forward,
nnU-Net v1 predict_3D sliding window, label ensemble / remap, Dice, connected-component and morphology statistics).
Every function cites the reference file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this package, and only as the checker or the
timed CPU baseline — the product path (``brainseg_b200``) never does.

Pinning status.  The reference ships no tests or golden vectors for this path, so the oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by ``oracle/make_golden.py`` (which imports
``/root/reference`` with stubs for its absent third-party imports) and committed under ``tests/golden/``:
  * Generic_UNet forward, state_dict key layout and FLOP structure  <- model_architecture/generic_UNet.py (imported)
  * label remap, Dice metrics, component / morphology statistics     <- convert_labels_to_brats.py,
    evaluate_segmentation.py, feature_extraction/{step3_multiplicity,step4_morphology,utils}.py (imported)
  * label-round ensemble LUT                                          <- run_brats2021_inference_singlethread.py:305
The nnU-Net v1 sliding-window internals (predict_3D, _get_gaussian, _compute_steps_for_sliding_window) live in the
un-vendored dependency ``Brats21_KAIST_MRI_Lab/nnunet`` (no pinned version in the reference; absent from
/root/reference): that part is restated from the published nnU-Net v1 algorithm and anchored
only on the call site run_brats2021_inference_singlethread.py:97-106 and the known-answer facts in
tests/golden/sliding_window.json — **parity unpinned** for that sub-part.
"""

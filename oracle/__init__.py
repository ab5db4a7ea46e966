"""CPU oracle for the top-down pose hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a plain numpy / torch-fp32 restatement of what
the reference (SamSamhuns/human_body_proportion_estimation) computes on the
hot path named in BASELINE.json.  It is the *checker*:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import it;
* nothing in ``human_body_proportion_estimation_b200/`` imports it, and the
  product path raises when its CUDA library is missing instead of coming here.

Pinning status (see DESIGN.md "Oracle"):
* decode / remap / gate / lengths / official NMS / legacy NMS / scale_coords /
  letterbox geometry: pinned against the reference's own functions executed in
  the authoring container (``oracle/ref_shim.py`` + ``tests/golden/make_golden.py``
  -> ``tests/golden/*.npz``).
* crop (``cv2.warpAffine`` fixed-point bilinear) and ``cv2.resize``: pinned
  against cv2 4.13 outputs stored in the golden files.
* ``tf.image.crop_and_resize`` geometry, EfficientDet person filter, HRNet
  network: **parity unpinned** -- the reference has no runnable implementation
  here (tensorflow / onnxruntime / the ONNX artifacts are absent); restated from
  the reference call sites and the public definitions.
"""

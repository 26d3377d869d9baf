"""CPU oracle for the Cap2Det proposal hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy fp32, one correctly-rounded op per
reference TF op; torch-CPU only for the Inception head convolutions) of the
reference algorithm.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``cap2det_b200/`` imports it and
the product path fails loudly when the CUDA library is missing.

Parity pin status (see DESIGN.md "Oracle"):

* pinned by the reference's own test vectors (``tests/golden/reference_vectors.json``,
  transcribed from ``core/box_utils_test.py``, ``core/utils_test.py``,
  ``models/label_extractor_test.py``, ``core/preprocess_test.py``):
  box_utils.{area,intersect,iou,flip_left_right,scale_to_new_size},
  masked_{maximum,minimum,sum,avg,sum_nd,avg_nd,softmax}, Groundtruth / ExactMatch /
  ExtendMatch extractors, parse_texts.
* **parity unpinned** (the reference has no test, and its numerics live in
  TensorFlow 1.15 / the un-vendored ``object_detection`` fork which cannot be
  installed here): crop_and_resize, max_pool, Mixed_5a-c head, MIDN, calc_oicr_loss,
  batch_multiclass_non_max_suppression, WordVectorMatch on real GloVe rows.
  For those the oracle restates the published TF kernel semantics (SURVEY.md
  Appendix A) and every decision is written next to the code.
"""

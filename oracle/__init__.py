"""CPU oracle for the ubdvss segment+CC hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker or as the timed CPU baseline.  The
product (``ubdvss_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Pinning status (see DESIGN.md "Oracle"):

* threshold + contour/box post-processing (``oracle.postproc``): PINNED.  The
  restatement is checked against the reference's own ``ModelRunner.predict`` /
  ``SegmapManager.postprocess`` / ``utils.get_contours_and_boxes`` executed in the
  build container (``tools/make_golden.py`` imports them from ``/root/reference``
  with the absent third-party modules stubbed) and against the committed fixtures
  in ``tests/golden/``.
* network forward (``oracle.net``), loss (``oracle.loss``) and Adam
  (``oracle.adam``): PARITY UNPINNED.  The arithmetic lives in TensorFlow-1.x /
  Keras-2.x (``requirements.txt:2,6``, unpinned, not installable here) and the
  reference ships no tests, golden vectors or weights.  The restatement follows
  the documented Keras/TF semantics; it is cross-checked by two independent
  implementations (NumPy shifted-slice loops vs torch-CPU ``conv2d``), analytic
  known-answer cases and an fp64 run.
"""

"""CPU oracle for the NBM audio hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker.
Nothing under ``birdsoundclassif_b200/`` imports it: the product path is the
CUDA library and fails loudly when that library is missing.

Parity status
-------------
* Post-processing (``postproc_oracle``): PINNED.  The restatement is checked
  against the reference's own functions (``nets_utils.nms``,
  ``bbox_reg_to_coord``, ``layers.ProposalLayer``, ``layers.FastRCNN`` tail,
  ``run_detection.merge_images``) executed in the build container; the
  resulting vectors are committed under ``tests/golden/`` together with
  ``oracle/make_golden.py``.
* Front-end (``frontend_oracle``): the reference's own ``File_Processor`` code
  is executed unmodified through ``ref_shims`` and the restatement is checked
  bit-for-bit against it, BUT the STFT arithmetic itself lives in the
  third-party ``librosa`` (un-pinned in the reference's requirements.txt:1,
  absent from this image, no tests or golden spectrograms in the reference).
  ``ref_shims.stft`` restates librosa's published algorithm (>=0.10 defaults).
  That part is therefore **parity unpinned** by the reference itself.
"""

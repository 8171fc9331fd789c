"""CPU oracle for the temporal-video-transformer hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``data-efficient-video-transformers_b200``) may import this package:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do, and there only as the checker or the timed CPU baseline.

The reference (ed-fish/data-efficient-video-transformers) has no arithmetic of its own on this path: every
FLOP is a ``torch.nn`` library call.  The oracle is therefore the same
``torch.nn`` composition the reference builds, restated with the hard-coded widths (2048 / 896) turned
into parameters so that BASELINE.json's configurations (d = 512 / 768) can be expressed:

  ``oracle.param``       the restatement; every class/function cites the reference file:line it follows
  ``oracle.shims``       stand-ins for packages the reference imports that are absent from this image
  ``oracle.ref_loader``  imports the UNMODIFIED reference modules from /root/reference (this container
                         only; used by ``make_golden.py`` to pin ``oracle.param`` against the reference)
  ``oracle.make_golden`` regenerates ``tests/golden/*.pt``

Parity status: the reference's own tests pin nothing on this path (src/tests holds no model test), so
the pin is "outputs of the reference itself run here": ``make_golden.py`` runs the verbatim reference
classes and ``oracle.param`` on the same seeds, asserts they agree, and freezes the reference outputs
as fixtures that ``tests/test_oracle_golden.py`` re-checks on every run.  Rows with no reference
implementation (queries-vs-keys cross-attention, the KL term) are pinned to torch functional calls
instead and are marked "torch-pinned" in DESIGN.md.
"""

"""CPU oracle for the VideoVanish per-frame pixel pipeline.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or as the timed CPU baseline.  ``videovanish_b200`` never imports
this package; it fails loudly when its CUDA library is missing.

Layout
------
``prepost.py``      rows A1-A9 of SURVEY.md section 8a (in-tree stages of
                    ``/root/reference/diffuerase.py:26-31`` and ``:69-114``),
                    restated with the same cv2 / scipy / numpy calls the
                    reference makes, plus closed-form numpy *models* of those
                    library calls (the spec the CUDA kernels implement).
``propagation.py``  row A10 (ProPainter flow-guided prior; un-vendored
                    upstream, PARITY UNPINNED by the reference).
``chunk_blend.py``  row A11 (builder-defined spec; PARITY UNPINNED).
``painter.py``      next row N3 (SAM2 colour painter; pinned by tests/golden/paint.npz).
``wrapper.py``      next row N4 (DiffuEraser wrapper read_mask + blurred compose; recalled
                    upstream, PARITY UNPINNED; its OpenCV primitives are pinned against cv2).
``reference_harness.py``  imports the UNMODIFIED reference module from
                    ``/root/reference`` with stub model packages; exists only
                    in the build container and is used to pin this oracle and
                    to generate ``tests/golden``.

Parity pinning status
---------------------
A1-A8: pinned against the reference's own code executed in the build
container (``tests/golden/*.npz`` were produced by
``tests/golden/make_golden.py`` through ``reference_harness``).
A9, A10, A11: "parity unpinned" - the reference tree holds no code, tests or
golden vectors for them.
"""

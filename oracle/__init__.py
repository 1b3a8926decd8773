"""CPU oracle for the nkb-classification hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The shipped package (``nkb_classification_b200``)
never imports this module and fails loudly when its CUDA library is missing.

Parity status: the reference (nkb-tech/nkb-classification) ships no tests,
golden vectors or fixtures for this path, so the oracle
is pinned against (a) ``cv2.resize`` 4.13 itself -- the third-party routine the
reference reaches through albumentations (``nkb_classification/dataset.py:102``)
-- and (b) the reference's own ``losses.py`` / ``metrics.py`` imported from
``/root/reference`` when the golden fixtures under ``tests/golden/`` were
generated (``tests/golden/make_golden.py``).  albumentations itself is not
installed anywhere in this image, so its glue arithmetic (Normalize,
LongestMaxSize rounding, PadIfNeeded centring) is restated from the published
1.3.x source and is "parity unpinned" at that one boundary.
"""

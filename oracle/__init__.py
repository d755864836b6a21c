"""ORACLE - TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference hot path (``MangaOcr(img) -> str``).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or
the timed CPU baseline - never from ``manga_ocr_b200`` (the product).

Parity status: the reference repository has NO tests, golden vectors or
fixtures for this path, and its arithmetic lives in
the un-vendored, unpinned pip packages ``manga-ocr`` -> ``transformers`` ->
``torch`` / ``Pillow``.  The oracle is therefore pinned against outputs of
those very libraries run in the build container (transformers 5.5.0, torch
2.11.0 CPU fp32, Pillow 12.2.0): ``oracle/make_golden.py`` generated the
fixtures under ``tests/golden/``.  What stays UNPINNED (no source offline) is the
~60-line ``manga_ocr.MangaOcr.__call__`` / ``post_process`` wrapper and
``jaconv.h2z``, restated from their published behaviour.
"""

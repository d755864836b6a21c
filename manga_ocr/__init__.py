"""Shim with the import name of the pip package the app uses: ``from manga_ocr import MangaOcr``
(reference/src/core/config.py:433) resolves to the B200-native engine with zero edits to the app."""
from manga_ocr_b200.ocr import MangaOcr  # noqa: F401

__all__ = ["MangaOcr"]

"""B200-native Manga-OCR recognition (``MangaOcr(img) -> str``): hand-written sm_100a CUDA behind
the C ABI of include/mocr_b200.h.  Importing this package makes no CUDA call."""
__version__ = "0.1.0"

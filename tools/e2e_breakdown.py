"""Where the end-to-end (host buffers in, strings out) time of one batch goes, beyond the resident-path time:
  python tools/e2e_breakdown.py [T=300] [B=64]   (B > 64: the mixed-size page crops)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from manga_ocr_b200 import crops as C, weights as W, text
from manga_ocr_b200.engine import Engine, _as_crop_array, RGB
T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=T)
crops = C.bubble_batch(B) if B <= 64 else C.page_batch(B, seed=1003)
vocab = text.Vocab.synthetic()

def t(f, n=10):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3

ids = eng.recognize(crops, RGB, T)[0]
print("bytes in", sum(c.nbytes for c in crops))
print("crop struct build   %.3f ms" % t(lambda: _as_crop_array(crops)))
def stage():
    eng.stage(crops); eng.sync()
print("stage + H2D + sync  %.3f ms" % t(stage))
def pre():
    eng.preprocess(); eng.sync()
print("preprocess          %.3f ms" % t(pre))
def enc():
    eng.encode(); eng.sync()
print("encode              %.3f ms" % t(enc))
def dec():
    eng.decode(T); eng.sync()
print("decode              %.3f ms" % t(dec, 5))
print("fetch ids           %.3f ms" % t(lambda: eng.fetch_ids()))
print("ids_to_text per row %.3f ms" % t(lambda: [text.ids_to_text(vocab, r) for r in ids]))
print("ids_to_texts batch %.3f ms" % t(lambda: text.ids_to_texts(vocab, ids)))
_, slow, cp, dots = vocab._fast_tables()
keep = ~vocab._special_mask[ids]
print("rows on the per-row path:", int((((slow[ids] & ~dots[ids]) | (cp[ids] == 0)) & keep).any(axis=1).sum()), "rows with dot rules:", int((dots[ids] & keep).any(axis=1).sum()))
print("recognize total     %.3f ms" % t(lambda: eng.recognize(crops, RGB, T), 5))
big = np.concatenate([c.reshape(-1) for c in crops]); dst = np.empty_like(big)
print("numpy memcpy same bytes %.3f ms" % t(lambda: np.copyto(dst, big)))

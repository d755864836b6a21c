"""Small end-to-end case for compute-sanitizer (every decoder program variant, taps, ragged crops)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine, TAP_ENCODER, TAP_LOGITS, TAP_PIXELS
T = 10
eng = Engine(W.random_init(0, eos_bias=3.0, gain=3.0), device=0, max_batch=5, max_length=T)
crops = C.page_batch(5, seed=3)
eng.set_taps(TAP_PIXELS | TAP_ENCODER | TAP_LOGITS)
for key, val, back in (("fuse_ln", 1, 1), ("fuse_ln", 0, 1), ("big_rows", 1, 96), ("kv_prefetch", 1, 0)):
    eng.set_option(key, val)
    ids, lens = eng.recognize(crops, max_length=T)
    print(key, val, lens.tolist())
    eng.set_option(key, back)
eng.set_option("gemm_pair", 1)
eng.stage(crops); eng.preprocess(); eng.encode(); eng.decode(T, forced_ids=np.zeros((5, T), np.int32) + 7)
print(eng.step_logits().shape, eng.encoder_hidden().shape)
eng.close()
print("done")

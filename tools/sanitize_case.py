"""Small end-to-end case for compute-sanitizer (every decoder program variant, slot refill, beam search, taps, ragged crops)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine, Region, TAP_ENCODER, TAP_LOGITS, TAP_PIXELS
T = 10
eng = Engine(W.random_init(0, eos_bias=3.7, gain=3.0), device=0, max_batch=6, max_length=T)
crops = C.page_batch(5, seed=3)
eng.set_taps(TAP_PIXELS | TAP_ENCODER | TAP_LOGITS)
for key, val, back in (("fuse_ln", 1, 1), ("fuse_ln", 0, 1), ("big_rows", 1, 112), ("kv_prefetch", 1, 0)):
    eng.set_option(key, val)
    ids, lens = eng.recognize(crops, max_length=T)
    print(key, val, lens.tolist())
    eng.set_option(key, back)
eng.stage(crops); eng.preprocess(); eng.encode(); eng.decode(T, forced_ids=np.zeros((5, T), np.int32) + 7)
print(eng.step_logits().shape, eng.encoder_hidden().shape)
eng.set_taps(0)
for slots, pipe in ((2, 0), (2, 1)):
    eng.set_option("slots", slots); eng.set_option("pipeline", pipe)
    ids, lens = eng.recognize(crops + crops[:3], max_length=T)     # 8 crops: two chunks of the 6-crop handle
    print("slots", slots, "pipeline", pipe, lens.tolist())
eng.set_option("slots", 0)
for dev in (1, 0):
    eng.set_option("beam_device", dev)
    ids, lens, scores = eng.recognize_beam(crops[:3], max_length=T, num_beams=2)
    print("beam device", dev, lens.tolist())
page, sels = C.page_with_selections(4, seed=9, height=900, width=700)
ids, lens = eng.recognize_regions(page, [Region.from_qt(r, p, o) for r, p, o in sels], max_length=T)
print("regions", lens.tolist())
eng.close()
print("done")

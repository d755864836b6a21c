"""Where the time of the cluster projection + LayerNorm kernel goes (build with NVCC_EXTRA=-DMOCR_PL_STAMPS)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
B, T = 64, 40
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=T)
eng.stage(C.bubble_batch(B)); eng.preprocess(); eng.encode()
eng.decode(T); eng.sync()
eng.set_option("decode_prof", 1)
eng.decode(T); eng.sync()
prof = eng.decode_profile(4096)
st = prof[3000:3008].astype(np.float64)
names = ["entry", "weights+consts issued", "dependency resolved", "MMAs issued (acc ready)", "tile reduced, local stats", "cluster phase 0 passed",
         "stats exchanged (cluster barrier)", "rows written"]
for i in range(1, 8):
    print(f"{names[i]:36s} +{(st[i]-st[i-1])/1e3:6.2f} us   (t = {(st[i]-st[0])/1e3:6.2f})")

"""Time the three decode modes on a GPU box:  python tools/decode_modes.py [B] [T]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=T)
eng.stage(C.bubble_batch(B)); eng.preprocess(); eng.encode()
ref = None
for mode, name in ((0, "v1 tcgen05 kernels + graph"), (1, "persistent cooperative"), (2, "stage kernels + graph")):
    eng.set_option("decode_mode", mode)
    for _ in range(2):
        eng.decode(T); eng.sync()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        eng.decode(T)
    eng.sync()
    dt = (time.perf_counter() - t0) / n
    ids, _ = eng.fetch_ids()
    if ref is None: ref = ids
    print(f"mode {mode} ({name}): {dt*1e3:.2f} ms per decode, {dt*1e6/eng.last_steps:.1f} us/step, steps {eng.last_steps}, ids equal to mode 0: {np.array_equal(ids, ref)} ({(ids==ref).mean():.4f})")

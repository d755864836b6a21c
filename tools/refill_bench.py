"""Slot refill variants on the ragged-length workload:  python tools/refill_bench.py [n_crops=512]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
crops = C.bubble_batch(n, seed=1002)
eng = Engine(W.random_init(0, eos_bias=4.2, gain=3.0), device=0, max_batch=n, max_length=300)
ref = None
for slots, pipe in ((64, 0), (64, 1), (64, 2), (128, 0), (32, 0), (0, 0)):
    eng.set_option("slots", slots); eng.set_option("pipeline", pipe)
    ids, lens = eng.recognize(crops)
    if ref is None: ref = ids
    t0 = time.perf_counter()
    for _ in range(3): eng.recognize(crops)
    dt = (time.perf_counter() - t0) / 3
    print(f"slots={slots} pipeline={pipe}: {dt*1e3:.1f} ms, {n/dt:.0f} crops/s, steps {eng.last_steps}, same ids {np.array_equal(ids, ref)}", flush=True)

import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
B, T = 64, 300
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=T)
eng.stage(C.bubble_batch(B)); eng.preprocess(); eng.encode()
opts = [dict(kv_evict_first=0), dict(kv_evict_first=1), dict(kv_evict_first=3), dict(kv_evict_first=2), dict(kv_evict_first=1), dict(kv_evict_first=3)]
for o in opts:
    for k, v in o.items(): eng.set_option(k, v)
    for _ in range(2):
        eng.decode(T); eng.sync()
    t0 = time.perf_counter(); n = 4
    for _ in range(n): eng.decode(T)
    eng.sync()
    dt = (time.perf_counter() - t0) / n
    print(f"{o}: {dt*1e3:.2f} ms, {dt*1e6/299:.1f} us/step", flush=True)

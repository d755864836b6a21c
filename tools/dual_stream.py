"""Experiment: one 64-row decode vs two concurrent 32-row decodes (two handles, two streams)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
T = 300
w = W.random_init(0)
crops = C.bubble_batch(64)
def prep(n, cr):
    e = Engine(w, device=0, max_batch=n, max_length=T)
    e.stage(cr); e.preprocess(); e.encode(); e.decode(T); e.sync()
    return e
e64 = prep(64, crops)
t0 = time.perf_counter()
for _ in range(3): e64.decode(T)
e64.sync()
print(f"1 x 64 rows: {(time.perf_counter()-t0)/3*1e3:.2f} ms per decode")
for parts in (2, 4):
    n = 64 // parts
    es = [prep(n, crops[i*n:(i+1)*n]) for i in range(parts)]
    def run(e):
        for _ in range(3): e.decode(T)
        e.sync()
    t0 = time.perf_counter()
    ts = [threading.Thread(target=run, args=(e,)) for e in es]
    for t in ts: t.start()
    for t in ts: t.join()
    print(f"{parts} x {n} rows concurrently: {(time.perf_counter()-t0)/3*1e3:.2f} ms per 64-row decode")
    e1 = es[0]
    t0 = time.perf_counter()
    for _ in range(3): e1.decode(T)
    e1.sync()
    print(f"   (one {n}-row decode alone: {(time.perf_counter()-t0)/3*1e3:.2f} ms)")
    for e in es: e.close()

"""Whole-path throughput vs batch size on one GPU (resident crops):  python tools/batch_sweep.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
T = 300
w = W.random_init(0)
for B in (16, 64, 128, 256, 512):
    eng = Engine(w, device=0, max_batch=B, max_length=T)
    eng.stage(C.bubble_batch(B))
    for _ in range(2):
        eng.run_resident(T); eng.sync()
    t0 = time.perf_counter(); n = 3
    for _ in range(n):
        eng.run_resident(T)
    eng.sync()
    dt = (time.perf_counter() - t0) / n
    print(f"B={B}: {dt*1e3:.1f} ms per batch, {B/dt:.0f} crops/s, {B*299/dt/1e3:.0f} k tok/s", flush=True)
    eng.close()

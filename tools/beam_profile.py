"""Device-resident beam mode only (for a launch list under ncu): python tools/beam_profile.py [T=300]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
eng = Engine(W.random_init(0), device=0, max_batch=64, max_length=T)
crops = C.bubble_batch(16)
for kv in sys.argv[2:]:
    k, v = kv.split("="); eng.set_option(k, int(v))
eng.recognize_beam(crops, max_length=T)
t0 = time.perf_counter()
ids, lens, scores = eng.recognize_beam(crops, max_length=T)
dt = time.perf_counter() - t0
print(f"beam 4 x 16 crops, T={T}: {dt*1e3:.1f} ms, steps {eng.last_steps}, {dt*1e6/max(eng.last_steps,1):.0f} us/step", flush=True)

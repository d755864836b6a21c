"""Beam mode throughput (not the headline): 16 crops x 4 beams per pass on a max_batch-64 handle, host bookkeeping vs device-resident."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
eng = Engine(W.random_init(0), device=0, max_batch=64, max_length=T)
crops = C.bubble_batch(16)
for dev, spg in ((0, 8), (1, 8), (1, 4), (1, 16)):
    eng.set_option("beam_device", dev); eng.set_option("beam_steps_per_graph", spg)
    eng.recognize_beam(crops, max_length=T)
    t0 = time.perf_counter()
    ids, lens, scores = eng.recognize_beam(crops, max_length=T)
    dt = time.perf_counter() - t0
    print(f"beam 4 x 16 crops, T={T}, device={dev} steps/graph={spg}: {dt*1e3:.1f} ms, {16/dt:.1f} crops/s, steps {eng.last_steps}, "
          f"{dt*1e6/max(eng.last_steps,1):.0f} us/step, mean len {lens.mean():.1f}", flush=True)
t0 = time.perf_counter()
eng.recognize(crops, max_length=T)
dt = time.perf_counter() - t0
print(f"greedy 16 crops: {dt*1e3:.1f} ms, {16/dt:.1f} crops/s")

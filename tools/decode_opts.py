import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
B, T = 64, 300
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=T)
eng.stage(C.bubble_batch(B)); eng.preprocess(); eng.encode()
eng.set_option("decode_mode", 2)
for graph, pdl, spg in ((0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 1), (1, 0, 13), (1, 1, 13), (1, 1, 26)):
    eng.set_option("use_graph", graph); eng.set_option("use_pdl", pdl); eng.set_option("steps_per_graph", spg)
    for _ in range(2):
        eng.decode(T); eng.sync()
    t0 = time.perf_counter(); n = 4
    for _ in range(n): eng.decode(T)
    eng.sync()
    dt = (time.perf_counter() - t0) / n
    print(f"graph={graph} pdl={pdl} steps/graph={spg}: {dt*1e3:.2f} ms, {dt*1e6/299:.1f} us/step", flush=True)

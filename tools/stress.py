"""Stability: many calls with varying batch sizes / entry points on one handle, and handle create/destroy cycles;
device memory in use must come back to where it started."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine, Region
w = W.random_init(0, gain=3.0, eos_bias=4.2)
free0 = torch.cuda.mem_get_info()[0]
rng = np.random.default_rng(0)
pool = C.page_batch(96, seed=1003)
page, sels = C.page_with_selections(24)
regions = [Region.from_qt(r, p, o) for r, p, o in sels]
t0 = time.time()
after = []
for cycle in range(5):
    eng = Engine(w, device=0, max_batch=32, max_length=48)
    used = []
    for it in range(60):
        n = int(rng.integers(1, 70))
        idx = rng.integers(0, len(pool), n)
        kind = it % 4
        if kind == 0:
            ids, lens = eng.recognize([pool[i] for i in idx])
        elif kind == 1:
            ids, lens = eng.recognize_regions(page, [regions[i % 24] for i in idx])
        elif kind == 2:
            ids, lens, _ = eng.recognize_beam([pool[i] for i in idx[:9]], num_beams=4)
        else:
            ids, lens = eng.recognize([pool[i] for i in idx], max_length=int(rng.integers(2, 49)))
        assert (ids[:, 0] == 2).all() and (lens >= 1).all()
        if it % 20 == 19:
            used.append(free0 - torch.cuda.mem_get_info()[0])
    eng.close()
    after.append(free0 - torch.cuda.mem_get_info()[0])
    print(f"cycle {cycle}: in-use while alive (MB) {[round(u / 2**20) for u in used]}, after close {round((free0 - torch.cuda.mem_get_info()[0]) / 2**20)} MB, {time.time() - t0:.0f} s", flush=True)
# (the first cycle loads the library's kernels and CUDA graphs' code: constant, not a leak)
growth = after[-1] - after[0]
print("after-close in-use per cycle (MB):", [round(a / 2**20) for a in after], "growth:", round(growth / 2**20, 1))
assert growth < 8 * 2**20
print("stress ok")

# ---- large batches: the large-batch program (> 112 rows), chunked staging (>= 256 crops), several passes (> max_batch), slot refill
big_pool = C.page_batch(160, seed=7)
after = []
for cycle in range(3):
    eng = Engine(w, device=0, max_batch=300, max_length=40)
    for it in range(12):
        n = int(rng.integers(100, 700))
        idx = rng.integers(0, len(big_pool), n)
        eng.set_option("slots", 64 if it % 3 == 2 else 0)
        ids, lens = eng.recognize([big_pool[i] for i in idx])
        assert ids.shape == (n, 40) and (ids[:, 0] == 2).all() and (lens >= 1).all()
    eng.close()
    after.append(free0 - torch.cuda.mem_get_info()[0])
    print(f"large cycle {cycle}: after close {round(after[-1] / 2**20)} MB in use, {time.time() - t0:.0f} s", flush=True)
assert after[-1] - after[0] < 8 * 2**20
print("ok")

"""Stage timeline of the persistent decoder (CTA 0, SM clocks) on a GPU box."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 120
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=T)
eng.stage(C.bubble_batch(B)); eng.preprocess(); eng.encode()
eng.set_option("decode_prof", 1)
for _ in range(2):
    eng.decode(T)
eng.sync()
prof = eng.decode_profile(4096)
NAMES = []
for l in range(2):
    NAMES += [f"L{l}.qkv", f"L{l}.self_attn", f"L{l}.self_out", f"L{l}.ln1", f"L{l}.cross_q", f"L{l}.cross_attn", f"L{l}.cross_out",
              f"L{l}.ln2", f"L{l}.fc1", f"L{l}.fc2", f"L{l}.ln3"]
NAMES += ["head_t", "head_ln", "vocab", "next_tok", "term_check"]
# stamps: [arrive(init)] then per stage: wait-exit, arrive ; term_check only has wait-exit
per_step = 2 * (len(NAMES) - 1) + 1
steps = (len(prof) - 1) // per_step
rows = []
for s in range(min(steps, T - 1)):
    base = 1 + s * per_step
    rows.append(prof[base: base + per_step])
rows = np.array(rows[5:], dtype=np.float64)      # skip the first steps
prev_arrive = np.concatenate([[np.nan], rows[0, 1:-1:2]])
work = rows[:, 1::2] - rows[:, 0:-1:2]           # wait-exit -> arrive
waitt = rows[:, 2::2] - rows[:, 1:-1:2]          # arrive -> next wait-exit (barrier latency + skew)
mhz = 1965.0
print(f"B={B} T={T} steps measured={len(rows)}; per-stage mean (us @ {mhz} MHz): work | barrier-after")
tot_w = tot_b = 0.0
for i, n in enumerate(NAMES[:-1]):
    w = work[:, i].mean() / mhz
    b = waitt[:, i].mean() / mhz
    tot_w += w; tot_b += b
    print(f"  {n:14s} {w:7.2f} | {b:7.2f}")
step_us = (rows[1:, 0] - rows[:-1, 0]).mean() / mhz
print(f"sum work {tot_w:.1f} us, sum barrier {tot_b:.1f} us, step {step_us:.1f} us")

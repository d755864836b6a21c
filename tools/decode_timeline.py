"""Timeline of the decoder's stage kernels (globaltimer of CTA 0) in the CUDA-graph mode."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = 40
opts = [a.split("=") for a in sys.argv[2:]]       # key=value engine options
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=T)
eng.stage(C.bubble_batch(B)); eng.preprocess(); eng.encode()
for k, v in opts:
    eng.set_option(k, int(v))
eng.decode(T); eng.sync()
eng.set_option("decode_prof", 1)
eng.decode(T); eng.sync()
prof = eng.decode_profile(4096)
n = int(prof[0]); rec = prof[1:1 + 4 * min(n, 1000)].reshape(-1, 4)
NAMES = {0: "gemm16", 1: "gemm32", 2: "gemm48", 3: "self_attn", 4: "cross_attn", 5: "ln", 6: "next", 7: "proj_ln"}
# a step ends with the next-token kernel (tag 6xx); the vocabulary GEMM runs on the tcgen05 kernel and leaves no record
ends = [i for i, r in enumerate(rec) if int(r[0]) // 100 == 6]
rec = rec[ends[7] + 1: ends[27] + 1]        # 20 steps, skipping the first ones
per = len(rec) / 20
t0 = rec[0, 1]
agg = {}
prev_done = None
for tag, te, tr, td in rec:
    name = NAMES[int(tag) // 100] + ("/k%d%s" % ((int(tag) % 100) // 10, "L" if int(tag) % 10 else "") if int(tag) // 100 < 3 else "")
    gap = (te - prev_done) if prev_done is not None else 0
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += (tr - te); a[2] += (td - tr); a[3] += gap
    prev_done = td
steps = 20.0
print(f"B={B} {sys.argv[2:]}: {len(rec)} records, {steps:.1f} steps, {(rec[-1,3]-rec[0,1])/steps/1e3:.1f} us/step")
print("stage            n/step  entry->ready  ready->done  prev_done->entry   (us, CTA 0)")
for k, (c, w, d, g) in agg.items():
    print(f"{k:16s} {c/steps:5.1f}   {w/c/1e3:8.2f}     {d/c/1e3:8.2f}     {g/c/1e3:8.2f}")

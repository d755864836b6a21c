"""us per launch of every named kernel (mocr_time_kernel), after one full decode."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
B, T = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 300
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=T)
eng.stage(C.bubble_batch(B)); eng.preprocess(); eng.encode()
for _ in range(2):
    eng.decode(T); eng.sync()
t0 = time.perf_counter(); n = 4
for _ in range(n): eng.decode(T)
eng.sync()
dt = (time.perf_counter() - t0) / n
print(f"decode {dt*1e3:.2f} ms, {dt*1e6/(T-1):.1f} us/step")
for k in ("dec_qkv", "dec_self_attn", "dec_self_out", "dec_ln", "dec_cross_attn", "dec_fc1", "dec_fc2", "dec_vocab"):
    ms, by, fl = eng.time_kernel(k, 50)
    print(f"{k:16s} {ms*1e3:7.2f} us")

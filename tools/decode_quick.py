"""Decode-only timing for a list of option sets:  python tools/decode_quick.py [B=64] [T=300] fuse_ln=1,kv_prefetch=1 fuse_ln=0,kv_prefetch=0 ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
args = sys.argv[1:]
B, T = 64, 300
while args and args[0].split("=")[0] in ("B", "T"):
    k, v = args.pop(0).split("=")
    if k == "B": B = int(v)
    else: T = int(v)
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=T)
eng.stage(C.bubble_batch(B)); eng.preprocess(); eng.encode()
for cfg in (args or [""]):
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("="); eng.set_option(k, int(v))
    for _ in range(2):
        eng.decode(T); eng.sync()
    t0 = time.perf_counter(); n = 4
    for _ in range(n): eng.decode(T)
    eng.sync()
    dt = (time.perf_counter() - t0) / n
    print(f"B={B} T={T} [{cfg}]: {dt*1e3:.2f} ms, {dt*1e6/(T-1):.1f} us/step", flush=True)

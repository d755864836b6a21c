"""Which side bounds the decoder's tcgen05 stage GEMMs at B rows?  Needs a -DMOCR_GEMM_DBG build (see tools/gemm_limits.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=300)
for kv in sys.argv[2:]:
    k, v = kv.split("="); eng.set_option(k, int(v))
eng.stage(C.bubble_batch(B)); eng.preprocess(); eng.encode(); eng.decode(40); eng.sync()
for name in ("dec_qkv", "dec_self_out", "dec_fc1", "dec_fc2", "dec_vocab"):
    row = []
    for dbg in (0, 1, 2, 3, 4, 0):
        eng.set_option("gemm_dbg", dbg)
        ms, by, fl = eng.time_kernel(name, 50)
        row.append(f"dbg{dbg} {ms*1e3:6.2f} us")
    print(f"B={B} {name:12s} " + " | ".join(row), flush=True)

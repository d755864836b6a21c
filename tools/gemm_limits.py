"""Which side bounds the encoder GEMMs?  Needs a build with -DMOCR_GEMM_DBG (NVCC_EXTRA="-DMOCR_GEMM_DBG" python -c "import __graft_entry__ as g; g.build()").
dbg 1: TMA loads stop after the first ring fill (MMA + epilogue speed); 2: no MMA issued (TMA + epilogue speed); 3: no epilogue work."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=16)
eng.stage(C.bubble_batch(B)); eng.preprocess(); eng.encode(); eng.sync()
for kv in sys.argv[2:]:
    k, v = kv.split("="); eng.set_option(k, int(v))
for name in ("enc_qkv", "enc_fc1", "enc_fc2", "enc_out"):
    row = []
    for dbg in (0, 1, 2, 3, 0):
        eng.set_option("gemm_dbg", dbg)
        ms, by, fl = eng.time_kernel(name, 30)
        row.append(f"dbg{dbg} {ms*1e3:7.2f} us ({fl/ms/1e9:6.0f} TF/s)")
    print(f"B={B} {name:8s} " + " | ".join(row), flush=True)
eng.set_option("gemm_dbg", 0)

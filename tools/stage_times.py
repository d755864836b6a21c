"""Each stage kernel of the decoder's per-token program timed alone (back-to-back launches, CUDA events) on the state a decode left:
python tools/stage_times.py [B=512] [T=300] [key=value engine options ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
args = sys.argv[1:]
B, T = 512, 300
while args and args[0].split("=")[0] in ("B", "T"):
    k, v = args.pop(0).split("=")
    if k == "B": B = int(v)
    else: T = int(v)
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=300)
for kv in args:
    k, v = kv.split("="); eng.set_option(k, int(v))
eng.stage(C.bubble_batch(B)); eng.preprocess(); eng.encode()
eng.decode(T); eng.sync()
total = 0.0
per_step = {"dec_qkv": 2, "dec_self_attn": 2, "dec_self_out": 2, "dec_ln": 7, "dec_cross_attn": 2, "dec_fc1": 2, "dec_fc2": 2, "dec_vocab": 1}
for name, n in per_step.items():
    try:
        ms, by, fl = eng.time_kernel(name, 50)
    except Exception as e:
        print(name, "n/a", e); continue
    total += ms * 1e3 * n
    print(f"{name:16s} {ms*1e3:8.2f} us  x{n}  {by/ms/1e6:8.1f} GB/s algorithmic")
print(f"B={B} keys={T-1}: sum over the step ~{total:.1f} us (cross-q, head transform not listed)")

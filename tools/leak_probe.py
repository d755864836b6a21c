"""Device memory after create / use / destroy cycles of a handle (a leak shows as a shrinking free figure):  python tools/leak_probe.py [full|create]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
w = W.random_init(0)
crops = C.bubble_batch(8)
torch.cuda.init(); free0 = torch.cuda.mem_get_info()[0]
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
for cycle in range(8):
    eng = Engine(w, device=0, max_batch=8, max_length=16)
    if mode != "create":
        eng.stage(crops); eng.preprocess(); eng.encode()
        if mode == "full":
            eng.set_option("use_graph", 1); eng.decode(16)
        elif mode == "nograph":
            eng.set_option("use_graph", 0); eng.decode(16)
        eng.sync()
    eng.close()
    print(mode, cycle, round((free0 - torch.cuda.mem_get_info()[0]) / 2**20), "MB", flush=True)

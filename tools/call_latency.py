"""Latency of MangaOcr(img) for a lone caller and for concurrent callers on one GPU, texts of ragged length (random-init
weights with an EOS bias: mean length ≈ 30 tokens):
python tools/call_latency.py [threads ...] [steps=K,option=value,... ...]   (steps = decode steps per session chunk = per CUDA graph, default 13;
the rest are engine options; MOCR_SESSION_PROF=1 prints where the dispatcher and mocr_session_add spend their time)"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from PIL import Image
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.ocr import MangaOcr
n = 1024
crops = [Image.fromarray(c) for c in C.bubble_batch(64, seed=1002)] * (n // 64)
w = W.random_init(0, eos_bias=4.2, gain=3.0)
devices = [int(x) for x in os.environ.get("MOCR_DEVICES", "0").split(",")]      # MOCR_DEVICES=0,1: one process, two GPUs
counts = [int(a) for a in sys.argv[1:] if "=" not in a] or [1, 4, 15, 50]
for cfg in [a for a in sys.argv[1:] if "=" in a] or ["steps=13"]:
    opts = dict(kv.split("=") for kv in cfg.split(","))
    steps = int(opts.pop("steps", 13))
    os.environ["MOCR_SESSION_STEPS"] = str(steps)
    rows = int(opts.pop("rows", 0))         # decoder rows of the session (0: 64)
    ocr = MangaOcr(weights=w, devices=devices, max_batch=64, max_length=300, warmup=True, slots=rows or None)
    for _ in range(50):                     # (the warm-up call's session may still be closing: options are refused until it has)
        try:
            for e in ocr.engines: e.set_option("steps_per_graph", steps)
            break
        except Exception:
            time.sleep(0.02)
    for k, v in opts.items():
        for e in ocr.engines: e.set_option(k, int(v))
    for threads in counts:
        for rep in range(2):
            it = iter(range(n)); lock = threading.Lock(); lat = []
            if ocr._session_prof is not None: ocr._session_prof.clear()
            def worker():
                while True:
                    with lock:
                        i = next(it, None)
                    if i is None: return
                    t0 = time.perf_counter(); ocr(crops[i]); lat.append(time.perf_counter() - t0)
            ts = [threading.Thread(target=worker) for _ in range(threads)]
            t0 = time.perf_counter()
            for t in ts: t.start()
            for t in ts: t.join()
            dt = time.perf_counter() - t0
            if rep:
                a = np.sort(np.array(lat)) * 1e3
                if ocr._session_prof:
                    print("   dispatcher phases (ms total, count):", {k: (round(v[0] * 1e3, 1), v[1]) for k, v in ocr._session_prof.items()}, f"wall {dt*1e3:.0f} ms")
                print(f"[{cfg}] {threads} callers: {n/dt:.0f} crops/s, latency ms p50 {a[len(a)//2]:.2f} p90 {a[int(len(a)*.9)]:.2f} max {a[-1]:.2f}", flush=True)
    ocr.close()

"""Latency of MangaOcr(img) for a lone caller and for a few concurrent callers on one GPU, texts of ragged length
(random-init weights with an EOS bias: mean length ≈ 30 tokens):  python tools/call_latency.py [threads ...]"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from PIL import Image
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.ocr import MangaOcr
n = 128
crops = [Image.fromarray(c) for c in C.bubble_batch(64, seed=1002)] * (n // 64)
ocr = MangaOcr(weights=W.random_init(0, eos_bias=4.2, gain=3.0), devices=[0], max_batch=64, max_length=300, warmup=True)
for threads in [int(a) for a in sys.argv[1:]] or [1, 4, 16]:
    for rep in range(2):
        it = iter(range(n)); lock = threading.Lock(); lat = []
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
            print(f"{threads} callers: {n/dt:.0f} crops/s, latency ms p50 {a[len(a)//2]:.2f} p90 {a[int(len(a)*.9)]:.2f} max {a[-1]:.2f}", flush=True)
ocr.close()

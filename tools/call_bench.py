"""50 threads calling MangaOcr(img) on one GPU (the `stream` leg's second half alone):
python tools/call_bench.py [key=value,... engine option sets]   (measured: capping the encoder GEMMs of an admission at 32 / 64 / 96 CTAs
changes nothing: 996-1007 crops/s)"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from PIL import Image
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.ocr import MangaOcr
n, threads = 384, 50
crops = [Image.fromarray(c) for c in C.tall_batch(64, seed=1005)] * (n // 64)
ocr = MangaOcr(weights=W.random_init(0), devices=[0], max_batch=64, max_length=300, warmup=True)
for cfg in (sys.argv[1:] or [""]):
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("="); ocr.engines[0].set_option(k, int(v))
    for rep in range(2):
        it = iter(range(n)); lock = threading.Lock()
        def worker():
            while True:
                with lock:
                    i = next(it, None)
                if i is None: return
                ocr(crops[i])
        ts = [threading.Thread(target=worker) for _ in range(threads)]
        t0 = time.perf_counter()
        for t in ts: t.start()
        for t in ts: t.join()
        dt = time.perf_counter() - t0
        if rep: print(f"[{cfg}] {n} crops, {threads} threads: {n/dt:.0f} crops/s", flush=True)
ocr.close()

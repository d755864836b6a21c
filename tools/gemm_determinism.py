"""Run-to-run determinism of the tcgen05 GEMM on the encoder's shapes (a race would show as differing outputs):
python tools/gemm_determinism.py [repeats=30]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from manga_ocr_b200 import weights as W
from manga_ocr_b200.engine import Engine
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
eng = Engine(W.random_init(0), device=0, max_batch=8, max_length=16)
for (M, N, K, bn, epi) in ((12608, 2304, 768, 256, 0), (12608, 3072, 768, 256, 1), (12608, 768, 3072, 256, 7), (512, 768, 3072, 0, 7)):
    rng = np.random.default_rng(M * 31 + N * 7 + K + bn + epi)
    A = rng.standard_normal((M, K), dtype=np.float32)
    Wt = rng.standard_normal((N, K), dtype=np.float32) * 0.05
    b = rng.standard_normal((N,), dtype=np.float32)
    R = rng.standard_normal((M, N), dtype=np.float32) if epi in (2, 7) else None
    first, bad = None, 0
    for i in range(reps):
        got, _ = eng.test_gemm(epi, bn, A, Wt, b, R)
        if first is None:
            first = got.copy()
        elif not np.array_equal(got, first):
            bad += 1
            d = np.argwhere(got != first)
            print(f"  run {i}: {len(d)} elements differ, first at {d[0].tolist()}: {got[tuple(d[0])]} vs {first[tuple(d[0])]}; rows {sorted(set((d[:,0]//128).tolist()))[:8]} (m-tiles) cols {sorted(set((d[:,1]//256).tolist()))[:8]} (n-tiles)", flush=True)
    print(f"M={M} N={N} K={K} bn={bn} epi={epi}: {reps} runs, {bad} differ from the first", flush=True)
eng.close()

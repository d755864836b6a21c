"""Developer smoke script for a GPU box: stage-by-stage parity of the CUDA path against the
oracle, with every stage isolated so one failure does not hide the rest.
  gpurun -- 'python tools/gpu_check.py > gpurun_out/check.log 2>&1'
"""
import json
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine, TAP_ENCODER, TAP_LOGITS, TAP_PIXELS

RES = {}
QUICK = "--quick" in sys.argv


def stage(name):
    def deco(fn):
        t = time.time()
        try:
            RES[name] = fn()
            print(f"[{name}] OK {RES[name]}  ({time.time() - t:.1f}s)", flush=True)
        except Exception:
            RES[name] = "FAILED"
            print(f"[{name}] FAILED\n{traceback.format_exc()}", flush=True)
        return fn
    return deco


def bf16_round(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).bfloat16().float().numpy()


weights = W.random_init(0)
T = 24
eng = Engine(weights, device=0, max_batch=8, max_length=T)
rng = np.random.default_rng(7)


@stage("gemm")
def _():
    out = {}
    for (M, N, K, bn, epi) in [(128, 256, 64, 32, 0), (128, 256, 128, 64, 0), (200, 768, 768, 128, 0), (333, 768, 256, 256, 0),
                               (64, 768, 768, 32, 2), (64, 3072, 768, 32, 1), (1000, 768, 3072, 192, 2), (8, 6144, 768, 64, 4),
                               (8, 768, 768, 32, 5), (1576, 2304, 768, 256, 0)]:
        A = rng.standard_normal((M, K), dtype=np.float32)
        Wt = rng.standard_normal((N, K), dtype=np.float32) * 0.05
        b = rng.standard_normal((N,), dtype=np.float32)
        R = rng.standard_normal((M, N), dtype=np.float32) if epi == 2 else None
        got, am = eng.test_gemm(epi, bn, A, Wt, b, R)
        ref = bf16_round(A).astype(np.float64) @ bf16_round(Wt).astype(np.float64).T + b
        if epi in (1, 5):
            import torch
            ref = torch.nn.functional.gelu(torch.from_numpy(ref)).numpy()
        if epi == 2:
            ref = ref + R
        err = float(np.abs(got - ref).max())
        tol = 2e-2 * max(1.0, float(np.abs(ref).max())) if epi in (0, 1) else 2e-3
        out[f"{M}x{N}x{K}/bn{bn}/epi{epi}"] = round(err, 6)
        assert err < tol, (M, N, K, bn, epi, err, tol)
        if epi == 4:
            assert np.array_equal(am, np.argmax(got, axis=1)), (am, np.argmax(got, axis=1))
    return out


@stage("enc_attention")
def _():
    import torch
    n = 2
    qkv = rng.standard_normal((n * 197, 2304), dtype=np.float32)
    qkv[:, :768] *= 0.125 * 2.0
    got = eng.test_encoder_attention(qkv)
    q = torch.from_numpy(bf16_round(qkv)).double().view(n, 197, 3, 12, 64)
    Q, K, V = q[:, :, 0].transpose(1, 2), q[:, :, 1].transpose(1, 2), q[:, :, 2].transpose(1, 2)
    P = torch.softmax(Q @ K.transpose(-1, -2), dim=-1)
    ref = (P @ V).transpose(1, 2).reshape(n * 197, 768).numpy()
    err = float(np.abs(got - ref).max())
    assert err < 3e-2, err
    return {"max_abs": err}


crops = C.bubble_batch(5, seed=1002) + [C.make_crop(rng, 224, 224), C.make_crop(rng, 900, 37, tint=True), C.make_crop(rng, 31, 500)]


@stage("preprocess")
def _():
    from oracle import preprocess_np as P
    eng.set_taps(TAP_PIXELS | TAP_ENCODER | TAP_LOGITS)
    eng.stage(crops)
    eng.preprocess()
    u8 = eng.pixels_u8()
    f32 = eng.pixel_values()
    bad = 0
    for i, c in enumerate(crops):
        ru8, rf = P.preprocess(c)
        bad += int((u8[i] != ru8).sum()) + int((f32[i].view(np.uint32) != rf[0].view(np.uint32)).sum())
    assert bad == 0, bad
    return {"mismatching_values": bad, "crops": len(crops)}


oracle = None


@stage("encoder")
def _():
    global oracle
    from oracle.reference_ocr import ReferenceMangaOcr
    from manga_ocr_b200.text import Vocab
    oracle = ReferenceMangaOcr(weights, Vocab.synthetic().tokens, max_length=T)
    eng.encode()
    got = eng.encoder_hidden()
    ref = oracle.encoder_hidden(crops)
    rel = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    mx = float(np.abs(got - ref).max())
    assert rel < 2e-2, (rel, mx)
    return {"rel_l2": rel, "max_abs": mx}


@stage("decode_teacher_forced")
def _():
    ids_ref, logits_ref = oracle.generate_batch(crops, max_length=T)
    ids_full = np.zeros((len(crops), T), np.int32)
    ids_full[:, : ids_ref.shape[1]] = ids_ref
    eng.decode(T, forced_ids=ids_full)
    got = eng.step_logits()[:, : logits_ref.shape[1]]
    ids_got, _ = eng.fetch_ids()
    mx = float(np.abs(got - logits_ref).max())
    rel = float(np.linalg.norm(got - logits_ref) / np.linalg.norm(logits_ref))
    top2 = np.sort(logits_ref, axis=-1)[..., -2:]
    margin = top2[..., 1] - top2[..., 0]
    am = ids_got[:, 1: 1 + logits_ref.shape[1]]
    mism = am != ids_ref[:, 1:]
    hard = int((mism & (margin > 6e-2)).sum())
    assert mx < 6e-2 and hard == 0, (mx, rel, hard)
    return {"logits_max_abs": mx, "rel_l2": rel, "argmax_mismatch": int(mism.sum()), "mismatch_above_margin": hard, "steps": int(logits_ref.shape[1])}


@stage("decode_free")
def _():
    ids_ref, _ = oracle.generate_batch(crops, max_length=T)
    for use_graph in (0, 1):
        eng.set_option("use_graph", use_graph)
        eng.decode(T)
        ids, lens = eng.fetch_ids()
        agree = float((ids[:, : ids_ref.shape[1]] == ids_ref).mean())
        print("  free-run graph=%d agree=%.4f lens=%s steps=%d" % (use_graph, agree, lens.tolist(), eng.last_steps))
    return {"token_agreement": agree}


@stage("recognize_e2e")
def _():
    ids, lens = eng.recognize(crops + crops, max_length=T)   # 16 crops through an 8-crop handle
    assert np.array_equal(ids[:8], ids[8:])
    return {"launches": eng.launch_count}


os.makedirs("gpurun_out", exist_ok=True)
json.dump(RES, open("gpurun_out/check.json", "w"), indent=1, default=str)
print(json.dumps(RES, default=str))
sys.exit(0 if all(v != "FAILED" for v in RES.values()) else 1)

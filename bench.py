#!/usr/bin/env python
"""Benchmark of the Manga-OCR recognition hot path (BASELINE.json metric: crops/sec, 224^2 crop,
greedy decode).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the whole path (fused preprocess -> ViT encoder -> cross-K/V projection ->
batched greedy decode to max_length 300) over one batch of 64 synthetic bubble crops per GPU
(BASELINE.json configs[1]; weak scaling: every rank takes its own 64 crops, no collective on the
math path, one final NCCL gather of the id rows in the e2e leg).  One JSON line on rank 0:
  value        crops/s, inputs already resident in HBM, CUDA-event timed, L2 flushed between steps
  e2e          crops/s through MangaOcr.recognize_batch with HOST crops (H2D + D2H + strings inside)
  roofline     dominant kernel: algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline the oracle (reference path on transformers, CPU fp32, batch 1) on a bounded sample
--impl reference times that CPU path as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "crops/sec (224^2 crop, greedy decode)"
UNIT = "crops/s"
BATCH = 64
MAX_LENGTH = 300
WORKLOAD = "configs[1]: manga-ocr-base arch, batch=64 synthetic bubble crops per GPU, bf16, greedy decode to max_length=300"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md).  NVML is polled in-process
    every 20 ms (the first sample is taken synchronously at start(), so even a 0.2 s region on an 8-GPU box is
    covered); `nvidia-smi -lms` - whose start-up alone can exceed a short region - is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip() != ""]
        self.index = int(ids[index]) if ids and index < len(ids) and ids[index].strip().isdigit() else index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []          # (sm_mhz, reason_bits)
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.t = None

    def _nvml_sample(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        try:
            bits = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:      # noqa: BLE001 - older bindings
            bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        self.samples.append((float(sm), int(bits)))

    def _nvml_loop(self):
        while not self.stop_flag.wait(0.02):
            try:
                self._nvml_sample()
            except Exception:  # noqa: BLE001
                return

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._nvml_sample()
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:      # noqa: BLE001 - no NVML bindings: nvidia-smi
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:      # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            try:
                self._nvml_sample()      # one more while the last kernels of the region are still in flight / just done
            except Exception:            # noqa: BLE001
                pass
            self.stop_flag.set()
            if self.t is not None:
                self.t.join(timeout=2)
            sm = [s[0] for s in self.samples]
            bits = 0
            for _, b in self.samples:
                bits |= b
            reasons = sorted(name for bit, name in self.BITS.items() if bits & bit)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------ CPU reference arm ---

def cpu_reference(crops, n_calls, max_length=None):
    """Reference path on the host CPU: PIL convert/resize + transformers generate, fp32, batch 1
    per call exactly like the app.  Returns (crops/s, tokens/s, threads)."""
    import torch
    from PIL import Image
    from manga_ocr_b200 import weights as W
    from manga_ocr_b200.text import Vocab
    from oracle.reference_ocr import ReferenceMangaOcr
    max_length = max_length or MAX_LENGTH
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ocr = cpu_reference.cache.get("ocr")
    if ocr is None:
        ocr = cpu_reference.cache["ocr"] = ReferenceMangaOcr(W.random_init(0), Vocab.synthetic().tokens, max_length=max_length)
    t0 = time.perf_counter()
    toks = 0
    for i in range(n_calls):
        img = Image.fromarray(crops[i % len(crops)])
        ids = ocr.generate_ids(img)
        toks += len(ids) - 1
    dt = time.perf_counter() - t0
    return n_calls / max(dt, 1e-9), toks / max(dt, 1e-9), torch.get_num_threads()


cpu_reference.cache = {}


def run_reference(args, rank):
    if rank != 0:
        return
    from manga_ocr_b200 import crops as C
    crops = C.bubble_batch(BATCH, seed=1002)
    per_step = 2
    cpu_reference(crops, 0)          # builds the model outside the timed region
    for _ in range(args.warmup):
        cpu_reference(crops, 1)
    t0 = time.perf_counter()
    done = 0
    toks_s = []
    for s in range(args.steps):
        cps, tps, cores = cpu_reference(crops[(s * per_step) % BATCH:] + crops, per_step)
        toks_s.append(tps)
        done += per_step
    dt = time.perf_counter() - t0
    value = done / dt
    sample = f"{per_step} crops per step (of the 64-crop batch), batch 1 per call, max_length {MAX_LENGTH}, fp32, torch CPU threads={cores}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "decode_tokens_per_s": float(np.mean(toks_s)) if toks_s else None,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ B200 arm ---

# algorithmic bytes of one decode step (bf16): the per-step weights once + per live crop its
# encoder K/V of both layers + the self-attention cache read (6144 B per cached token) + the 6144 B it appends
DEC_WEIGHT_BYTES = 43_716_096
CROSS_KV_BYTES = 1_210_368
SELF_KV_BYTES_PER_TOKEN = 6144
ENC_FLOPS_PER_CROP = 36.056e9       # 35.126 GFLOP encoder + 0.930 GFLOP cross-K/V projection


def decode_algorithmic_bytes(rows, steps):
    """Sum over token steps t = 1..steps of the section-8(d) bytes for `rows` live crops."""
    t = np.arange(1, steps + 1, dtype=np.float64)
    return float(np.sum(DEC_WEIGHT_BYTES + rows * (CROSS_KV_BYTES + SELF_KV_BYTES_PER_TOKEN * t + SELF_KV_BYTES_PER_TOKEN)))


NCU_KERNEL_OF = {"dec_qkv": "pd_gemm_kernel<16", "dec_fc2": "pd_gemm_kernel<16", "dec_fc1": "pd_gemm_kernel<32", "dec_self_out": "pd_proj_ln_kernel",
                 "dec_ln": "pd_ln_kernel", "dec_self_attn": "pd_attention_kernel<1>", "dec_cross_attn": "pd_attention_kernel<0>"}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each decoder stage kernel, from the committed ncu summary
    (profiles/r2_ncu_traffic.json, written from one `ncu --set full` capture; its command is inside).  {} when absent."""
    fp = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        with open(fp) as f:
            d = json.load(f)
        return {row["kernel"]: float(row["dram_bytes_per_launch"]) for row in d.get("kernels", [])}, d.get("command")
    except (OSError, ValueError, KeyError):
        return {}, None


def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from manga_ocr_b200 import crops as C, weights as W
    from manga_ocr_b200.engine import RGB
    from manga_ocr_b200.ocr import MangaOcr
    from manga_ocr_b200.splitter import gather_ids, shard_bounds
    from manga_ocr_b200.text import ids_to_texts

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 arm has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")        # host-side barriers while one rank drives several GPUs in-process
    legs = set(args.legs.split(",")) if args.legs != "all" else {"page512", "tall64", "stream", "regions", "ragged", "cpu"}
    weights = W.random_init(0)
    ocr = MangaOcr(weights=weights, devices=[local_rank], max_batch=BATCH, max_length=MAX_LENGTH, warmup=False)
    eng = ocr.engines[0]
    crops = C.bubble_batch(BATCH, seed=1002 + 7919 * rank)
    in_bytes = int(sum(c.nbytes for c in crops))
    stream = torch.cuda.ExternalStream(eng.stream, device=local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")   # > 126 MB L2
    peaks = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        if world > 1:
            dist.barrier(group=cpu_group)

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.add_(1)

    def max_over_ranks(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    # ---- resident leg: crops staged once, the device path timed with CUDA events
    eng.stage(crops, RGB)
    for _ in range(args.warmup):
        eng.run_resident(MAX_LENGTH)
    eng.sync()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = eng.launch_count
    evs = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush_l2()
        e0 = ev()
        eng.run_resident(MAX_LENGTH)
        evs.append((e0, ev()))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = eng.launch_count - launches0
    ms = sum(a.elapsed_time(b) for a, b in evs)
    steps_decoded = eng.last_steps

    # ---- e2e leg: host crops in, strings out, through the public batch API
    for _ in range(max(1, min(args.warmup, 2))):
        ocr.recognize_batch(crops)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if world == 1:
            texts = ocr.recognize_batch(crops)   # the public call: host uint8 crops -> strings
        else:
            ids = ocr.recognize_ids(crops)       # the same call in its two halves, to have the id rows for the gather
            texts = ids_to_texts(ocr.vocab, ids)
            gather_ids(ids, BATCH * world)       # final result gather: NCCL all_gather of [64, 300] int32 per rank
        assert len(texts) == BATCH
    barrier()
    e2e_s = time.perf_counter() - t0
    ms_max, e2e_ms_max = max_over_ranks(ms, e2e_s * 1e3)

    # ---- phases of one step, per-kernel table, roofline (rank 0; CUDA events on the engine's stream)
    roof, phases, kernels = None, None, []
    if rank == 0:
        ph = []
        for _ in range(3):
            flush_l2()
            e0 = ev(); eng.preprocess(); e1 = ev(); eng.encode(); e2 = ev(); eng.decode(MAX_LENGTH); e3 = ev()
            eng.sync()
            ph.append((e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)))
        steps_run = max(eng.last_steps, 1)
        phases = {"preprocess_ms": min(p[0] for p in ph), "encode_ms": min(p[1] for p in ph), "decode_ms": min(p[2] for p in ph),
                  "decode_us_per_token_step": 1e3 * min(p[2] for p in ph) / steps_run}
        phases["encoder_tflops"] = BATCH * ENC_FLOPS_PER_CROP / (phases["encode_ms"] * 1e-3) / 1e12
        phases["encoder_frac_of_sustained_bf16_peak"] = phases["encoder_tflops"] / peaks["bf16_tflops_sustained"]
        dec_bytes = decode_algorithmic_bytes(BATCH, steps_run)
        dec_gbs = dec_bytes / (phases["decode_ms"] * 1e-3) / 1e9
        floor_ms = 1e3 * (dec_bytes / (peaks["hbm_gbs"] * 1e9) + BATCH * ENC_FLOPS_PER_CROP / (peaks["bf16_tflops_sustained"] * 1e12))
        step_ms = ms_max / args.steps
        # launches of each stage kernel per token step (decode_stages.cuh: pd_build_program) or per encoder pass
        per_step = {"enc_ln": 25, "enc_qkv": 12, "enc_attn": 12, "enc_out": 12, "enc_fc1": 12, "enc_fc2": 12, "dec_qkv": 2, "dec_self_attn": 2,
                    "dec_self_out": 5, "dec_ln": 2, "dec_cross_attn": 2, "dec_cross_q": 2, "dec_fc1": 2, "dec_fc2": 2, "dec_vocab": 1}
        for name in ("enc_ln", "enc_qkv", "enc_attn", "enc_out", "enc_fc1", "enc_fc2", "dec_qkv", "dec_self_attn", "dec_self_out", "dec_ln",
                     "dec_cross_attn", "dec_fc1", "dec_fc2", "dec_vocab"):
            try:
                k_ms, k_bytes, k_flops = eng.time_kernel(name, 50)
            except Exception as e:      # noqa: BLE001
                kernels.append({"kernel": name, "error": str(e)})
                continue
            n_launch = per_step[name] * (steps_run if name.startswith("dec_") else 1)
            kernels.append({"kernel": name, "us_per_launch": 1e3 * k_ms, "launches_per_step": n_launch,
                            "share_of_step": k_ms * n_launch / step_ms, "algorithmic_bytes": k_bytes,
                            "GBps": k_bytes / (k_ms * 1e-3) / 1e9, "hbm_frac": k_bytes / (k_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                            "TFLOPs": k_flops / (k_ms * 1e-3) / 1e12})
        # self-attention time grows with the cache: the row above is the worst case (t = %d keys); its share of the step uses the
        # mean of that and the same kernel on a 20-token decode
        try:
            eng.decode(20)
            sa_short = eng.time_kernel("dec_self_attn", 50)[0]
            eng.decode(MAX_LENGTH)
            for r in kernels:
                if r.get("kernel") == "dec_self_attn":
                    r["us_per_launch_short_cache"] = 1e3 * sa_short
                    r["share_of_step"] = 0.5 * (r["us_per_launch"] * 1e-3 + sa_short) * r["launches_per_step"] / step_ms
        except Exception:      # noqa: BLE001
            pass
        kernels.sort(key=lambda r: -r.get("share_of_step", 0.0))      # dominant by time first
        top = next((r for r in kernels if "error" not in r), None)
        ncu_bytes, ncu_cmd = ncu_traffic()
        step_traffic = 0.0
        for r in kernels:
            pat = NCU_KERNEL_OF.get(r.get("kernel", ""))
            hit = next((v for k, v in ncu_bytes.items() if pat and pat in k), None)
            r["dram_bytes_per_launch_ncu"] = hit
            if hit is not None and r["kernel"].startswith("dec_"):
                step_traffic += hit * per_step[r["kernel"]]
        traffic = step_traffic if step_traffic > 0 else None
        traffic_src = None
        if traffic is not None:
            traffic += 2.0 * 6144 * 768 + 2 * 2.0 * 768 * 768    # + vocabulary projection and the two cross-q weights (not in the capture): algorithmic
            traffic_src = {"file": "profiles/r2_ncu_traffic.json", "command": ncu_cmd,
                           "what": "DRAM bytes of ONE token step = sum over the stage kernels of (cold-cache bytes per launch x launches per step), "
                                   "captured at ~20 cached tokens; compare with algorithmic_bytes_per_step_t20"}
        roof = {
            "bound": "hbm", "kernel": "decode token step: all stage kernels of the %d steps (%.0f %% of the step's time)" % (steps_run, 100 * phases["decode_ms"] / step_ms),
            "achieved": dec_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": dec_gbs / peaks["hbm_gbs"],
            "algorithmic_bytes": dec_bytes, "algorithmic_bytes_per_step_t20": DEC_WEIGHT_BYTES + BATCH * (CROSS_KV_BYTES + SELF_KV_BYTES_PER_TOKEN * 21),
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks["source"],
            "step_floor_ms": floor_ms, "step_frac_of_combined_roofline": floor_ms / step_ms,
            "dominant_kernel": ({k: top.get(k) for k in ("kernel", "us_per_launch", "launches_per_step", "share_of_step", "algorithmic_bytes", "GBps", "hbm_frac",
                                                           "dram_bytes_per_launch_ncu")} if top else None),
            "note": "frac = SURVEY 8(d) bytes of the whole decode / its CUDA-event time in the L2-flushed step / measured HBM peak; "
                    "step_frac_of_combined_roofline adds the encoder's tensor floor. The per-kernel rows (`kernels`, dominant by time first) are "
                    "50 back-to-back launches of one stage on the state the decode left (programmatic dependent launch as in the loop): L2-WARM, "
                    "so their GB/s are L2 + HBM delivery rates, not DRAM bandwidth"}

    extra = {}
    ocr.close()
    del ocr, eng
    torch.cuda.empty_cache()

    # ---- BASELINE configs[2]: one page batch of 512 mixed-size crops, STRONG-scaled over the ranks (host crops in, strings out)
    if "page512" in legs:
        total = args.page_crops
        lo, hi = shard_bounds(total, world, rank)
        page = C.page_batch(total, seed=1003)[lo:hi]
        pocr = MangaOcr(weights=weights, devices=[local_rank], max_batch=max(hi - lo, 1), max_length=MAX_LENGTH, warmup=False)
        pocr.recognize_batch(page[: min(len(page), 8)])
        pocr.recognize_batch(page)
        barrier()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            ids = pocr.recognize_ids(page)
            ids_to_texts(pocr.vocab, ids)
            if world > 1:
                gather_ids(ids, total)
        barrier()
        dt = (time.perf_counter() - t0) / reps
        # device-only share of the same work (resident crops)
        pe = pocr.engines[0]
        pe.stage(page, RGB)
        pstream = torch.cuda.ExternalStream(pe.stream, device=local_rank)
        a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
        pe.run_resident(MAX_LENGTH); pe.sync()     # untimed: the one-piece encoder graph of this batch size (the e2e path stages and encodes in chunks)
        a0.record(pstream); pe.run_resident(MAX_LENGTH); a1.record(pstream); pe.sync()
        dt_max, dev_ms = max_over_ranks(dt, a0.elapsed_time(a1))
        extra["page512"] = {"workload": "configs[2]: 512 mixed-size crops (seed 1003), strong scaling", "crops": total, "n_gpus": world,
                            "crops_per_rank": hi - lo, "e2e_crops_per_s": total / dt_max, "e2e_ms": dt_max * 1e3,
                            "resident_crops_per_s": total / (dev_ms * 1e-3), "decode_steps_run": pe.last_steps, "h2d_bytes": int(sum(c.nbytes for c in page)),
                            "api": "MangaOcr.recognize_ids + ids_to_texts per rank, NCCL all_gather of the id rows"}
        pocr.close()
        del pocr, pe

    # ---- BASELINE configs[3]: tall vertical-text crops, 299 decode steps (decoder / KV-cache bandwidth stress), weak scaling
    if "tall64" in legs:
        tall = C.tall_batch(BATCH, seed=1004 + 7919 * rank)
        tocr = MangaOcr(weights=weights, devices=[local_rank], max_batch=BATCH, max_length=MAX_LENGTH, warmup=False)
        te = tocr.engines[0]
        tstream = torch.cuda.ExternalStream(te.stream, device=local_rank)
        te.stage(tall, RGB)
        te.run_resident(MAX_LENGTH); te.sync()
        te.preprocess(); te.encode(); te.sync()
        barrier()
        best = None
        for _ in range(3):
            with torch.cuda.stream(tstream):
                flush.add_(1)
            a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
            a0.record(tstream); te.decode(MAX_LENGTH); a1.record(tstream); te.sync()
            best = a0.elapsed_time(a1) if best is None else min(best, a0.elapsed_time(a1))
        t0 = time.perf_counter()
        tocr.recognize_batch(tall)
        e2e_tall = time.perf_counter() - t0
        (dec_ms, e2e_tall) = max_over_ranks(best, e2e_tall)
        steps_t = max(te.last_steps, 1)
        tb = decode_algorithmic_bytes(BATCH, steps_t)
        extra["tall64"] = {"workload": "configs[3]: 64 tall crops per GPU (W 40-120, H 600-1600), max_length 300, weak scaling", "n_gpus": world,
                           "decode_ms": dec_ms, "decode_us_per_token_step": 1e3 * dec_ms / steps_t,
                           "decode_tokens_per_s": BATCH * world * steps_t / (dec_ms * 1e-3), "decode_hbm_GBps_per_gpu": tb / (dec_ms * 1e-3) / 1e9,
                           "decode_hbm_frac": tb / (dec_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "e2e_crops_per_s": BATCH * world / e2e_tall,
                           "h2d_bytes_per_gpu": int(sum(c.nbytes for c in tall))}
        tocr.close()
        del tocr, te

    # ---- BASELINE configs[4]: the crop stream through ONE MangaOcr(devices=[0..N-1]) process (rank 0 drives every GPU of the
    #      job; the other ranks have released theirs and wait on a host-side barrier): 50 caller threads on __call__ (the app's
    #      worker model, reference/src/ui/main_window.py:4317-4327) and the same crops through recognize_batch
    del flush
    torch.cuda.empty_cache()
    host_barrier()
    if "stream" in legs and rank == 0:
      try:
          from PIL import Image
          n_stream = args.stream_crops if args.stream_crops > 0 else 512 * world
          per_gpu = -(-n_stream // world)
          scrops = C.page_batch(n_stream, seed=1005)
          socr = MangaOcr(weights=weights, devices=list(range(world)), max_batch=min(512, per_gpu), max_length=MAX_LENGTH, warmup=False)
          socr.recognize_batch(scrops[: 8 * world])
          socr.recognize_batch(scrops)                  # warm: graphs, arenas, tables
          t0 = time.perf_counter()
          texts = socr.recognize_batch(scrops)
          dt_batch = time.perf_counter() - t0
          assert len(texts) == n_stream
          imgs = [Image.fromarray(c) for c in scrops]
          n_threads = 50
          out = [None] * n_stream
          nxt = [0]
          lock = threading.Lock()

          def caller():
              while True:
                  with lock:
                      i = nxt[0]
                      nxt[0] += 1
                  if i >= n_stream:
                      return
                  out[i] = socr(imgs[i])

          def run_calls():
              ts = [threading.Thread(target=caller) for _ in range(n_threads)]
              t0 = time.perf_counter()
              for t in ts:
                  t.start()
              for t in ts:
                  t.join()
              return time.perf_counter() - t0

          nxt[0] = max(0, n_stream - 64 * world)        # warm (like the batch call above): the session's graphs, the encoder graphs
          run_calls()                                   # of small admissions, the dispatcher threads - 64 crops per GPU, untimed
          nxt[0] = 0
          dt_call = run_calls()
          same = sum(a == b for a, b in zip(out, texts))
          same1 = sum(a[:1] == b[:1] for a, b in zip(out, texts))
          extra["stream"] = {"workload": "configs[4]: crop stream (config-3 distribution, seed 1005), one process driving every GPU", "crops": n_stream,
                             "in_process_gpus": world, "recognize_batch_crops_per_s": n_stream / dt_batch, "recognize_batch_s": dt_batch,
                             "call_threads": n_threads, "call_crops_per_s": n_stream / dt_call, "call_s": dt_call,
                             "call_first_char_equal_to_batch": same1 / n_stream, "call_strings_equal_to_batch": same / n_stream, "h2d_bytes": int(sum(c.nbytes for c in scrops)),
                             "note": "__call__ blocks its caller until that crop is decoded: 50 threads bound the crops in flight to 50 "
                                     "(<= 50 / GPUs per batch), whatever the engine could take. The batch call decodes >112 rows per GPU on the large-batch program "
                                     "(tcgen05 GEMMs), the small batches of __call__ on the mma.sync program: with random-init weights (top-2 margins "
                                     "of ~1e-2) the two roundings part at a near-tie somewhere in 299 tokens for most crops, hence the low "
                                     "whole-string agreement; parity of each program with the oracle is what tests/ pins"}
          socr.close()
          del socr
      except Exception as e:      # noqa: BLE001 - e.g. the launcher restricted this rank to one visible GPU
        extra["stream"] = {"error": f"{type(e).__name__}: {e}"}
    host_barrier()

    # ---- region staging: 64 selections of one page, crop + polygon composite + rotation on the device
    regions_info = None
    ocr = eng = None
    if rank == 0 and world == 1 and ({"regions", "ragged"} & legs):
        ocr = MangaOcr(weights=weights, devices=[local_rank], max_batch=BATCH, max_length=MAX_LENGTH, warmup=False)
        eng = ocr.engines[0]
    if rank == 0 and world == 1 and "regions" in legs:
        from manga_ocr_b200.engine import Region
        page, sels = C.page_with_selections(BATCH)
        regions = [Region.from_qt(rect, poly, orient) for rect, poly, orient in sels]
        ocr.recognize_regions(page, regions)
        t0 = time.perf_counter()
        for _ in range(2):
            ocr.recognize_regions(page, regions)
        reg_s = (time.perf_counter() - t0) / 2
        def timed(f, n=5):
            f()
            t0 = time.perf_counter()
            for _ in range(n):
                f()
            return (time.perf_counter() - t0) / n * 1e3
        def dev_stage():
            eng.stage_regions(page, regions)
            eng.sync()
        regions_info = {"selections": BATCH, "page": list(page.shape), "with_polygon": sum(r.polygon is not None for r in regions),
                        "rotated": sum(r.rotate != 0 for r in regions), "e2e_crops_per_s": BATCH / reg_s,
                        "device_staging_ms": timed(dev_stage), "api": "MangaOcr.recognize_regions (page + selections -> strings)"}
        if "cpu" in legs:
            try:
                from oracle.make_golden_staging import reference_stage     # the reference's own PIL / cv2 call sequence
                host_ms = timed(lambda: [reference_stage(page, r.box, r.polygon, r.rotate) for r in regions])
                regions_info["host_staging_ms_reference_libs"] = host_ms
            except Exception as e:      # noqa: BLE001 - cv2 missing on this box
                regions_info["host_staging_ms_reference_libs"] = None
                regions_info["host_staging_note"] = f"not measured: {e}"

    # ---- realistic length mix (SURVEY.md 8d "Weights": documented variant, NOT the headline): same architecture with
    #      cls.predictions.bias[3] raised so that rows emit EOS early
    ragged = None
    if rank == 0 and world == 1 and "ragged" in legs:
        from manga_ocr_b200.engine import Engine
        n_rag = args.ragged_crops
        rag = C.bubble_batch(n_rag, seed=1002)
        w_rag = W.random_init(0, eos_bias=args.eos_bias, gain=3.0)
        # (a) the reference's scheme: batches of 64, finished rows padded until the longest row of the batch ends
        eng2 = Engine(w_rag, device=local_rank, max_batch=BATCH, max_length=MAX_LENGTH)
        ids2, lens2 = eng2.recognize(rag, RGB, MAX_LENGTH)
        t0 = time.perf_counter()
        eng2.recognize(rag, RGB, MAX_LENGTH)
        dt_pad = time.perf_counter() - t0
        steps_pad = int(eng2.last_steps)
        eng2.close()
        # (b) in-flight slot refill: all crops encoded, 64 decoder rows, a row that finishes takes the next waiting crop
        eng3 = Engine(w_rag, device=local_rank, max_batch=n_rag, max_length=MAX_LENGTH)
        eng3.set_option("slots", BATCH)
        ids3, lens3 = eng3.recognize(rag, RGB, MAX_LENGTH)
        t0 = time.perf_counter()
        for _ in range(3):
            eng3.recognize(rag, RGB, MAX_LENGTH)
        dt = (time.perf_counter() - t0) / 3
        ragged = {"eos_bias": args.eos_bias, "gain": 3.0, "crops": n_rag, "decoder_rows": BATCH, "mean_len": float(lens3.mean()),
                  "max_len": int(lens3.max()), "tokens": int(lens3.sum() - n_rag), "decode_steps_run": int(eng3.last_steps),
                  "e2e_ids_crops_per_s": n_rag / dt, "e2e_tokens_per_s": float(lens3.sum() - n_rag) / dt,
                  "padded_batches_of_64": {"e2e_ids_crops_per_s": n_rag / dt_pad, "decode_steps_last_batch": steps_pad},
                  "ids_equal_to_padded_scheme": bool(np.array_equal(ids2, ids3) and np.array_equal(lens2, lens3)),
                  "note": "host crops in, ids out, through mocr_recognize; slot refill (option slots=64) vs the reference's padded batches"}
        eng3.close()
        # (c) the app's own call shape on the same crops: 50 threads calling MangaOcr(img), one crop each; the dispatcher admits
        #     them into the running decode session (main_window.py:608-611, 4317-4327)
        try:
            from PIL import Image
            imgs = [Image.fromarray(c) for c in rag]
            ocr_rag = MangaOcr(weights=w_rag, devices=[local_rank], max_batch=BATCH, max_length=MAX_LENGTH, warmup=True)
            best = None
            for rep in range(3):                           # (the first pass warms the resampling tables of the crop sizes)
                it = iter(range(n_rag))
                lk = threading.Lock()
                lat = []

                def caller():
                    while True:
                        with lk:
                            i = next(it, None)
                        if i is None:
                            return
                        t1 = time.perf_counter()
                        ocr_rag(imgs[i])
                        lat.append(time.perf_counter() - t1)
                ts = [threading.Thread(target=caller) for _ in range(50)]
                t0 = time.perf_counter()
                for t in ts:
                    t.start()
                for t in ts:
                    t.join()
                dt_call = time.perf_counter() - t0
                a = np.sort(np.array(lat)) * 1e3
                if rep and (best is None or dt_call < best[0]):
                    best = (dt_call, float(a[len(a) // 2]), float(a[int(len(a) * 0.9)]), float(a[-1]))
            ocr_rag.close()
            ragged["call_50_threads"] = {"crops_per_s": n_rag / best[0], "latency_ms_p50": best[1], "latency_ms_p90": best[2],
                                         "latency_ms_max": best[3], "api": "MangaOcr.__call__ (PIL image -> str), admission into the "
                                         "running decode session, 64 decoder rows; best of 2 passes over the crops"}
        except Exception as e:      # noqa: BLE001 - this leg must not take the headline line down
            ragged["call_50_threads"] = {"error": repr(e)}
    if ocr is not None:
        ocr.close()

    cpu = None
    if rank == 0 and world == 1 and "cpu" in legs and not args.no_cpu_baseline:
        n_calls = 8
        cpu_reference(crops, 1)
        cps, tps, cores = cpu_reference(crops, n_calls)
        cpu = {"value": cps, "unit": UNIT, "cores": cores, "kind": "port", "decode_tokens_per_s": tps,
               "sample": f"first {n_calls} crops of the batch, batch 1 per call, max_length {MAX_LENGTH}, fp32"}

    if rank == 0:
        total_crops = BATCH * world * args.steps
        value = total_crops / (ms_max * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "crops_per_gpu": BATCH, "max_length": MAX_LENGTH, "decode_steps_run": steps_decoded,
                       "weights": "random-init manga-ocr-base (numpy PCG64 seed 0)", "l2": "flushed between timed steps (256 MiB write)",
                       "parallelism": f"dp{world} (crops sharded, no collective on the math path)"},
            "decode_tokens_per_s": BATCH * world * steps_decoded * args.steps / (ms_max * 1e-3),
            "e2e": {"value": total_crops / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": in_bytes + 40 * BATCH,
                    "d2h_bytes_per_step": BATCH * MAX_LENGTH * 4 + BATCH * 4, "api": "MangaOcr.recognize_batch (host uint8 crops -> strings)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "phases": phases, "kernels": kernels,
            "configs": extra, "regions": regions_info, "ragged_lengths": ragged, "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--legs", default="all", help="comma list of the extra legs to run: page512,tall64,stream,regions,ragged,cpu (default all; 'none' = headline only)")
    ap.add_argument("--page-crops", type=int, default=512, help="crops of the strong-scaled page batch (BASELINE configs[2])")
    ap.add_argument("--stream-crops", type=int, default=0, help="crops of the stream leg (BASELINE configs[4]: 4096 on 8 GPUs); 0 = 512 per GPU")
    ap.add_argument("--ragged-crops", type=int, default=512, help="crops of the realistic-length leg")
    ap.add_argument("--eos-bias", type=float, default=4.2, help="EOS bias of the realistic-length leg's weights")
    ap.add_argument("--max-length", type=int, default=300, help="profiling only; the metric is defined at 300")
    args = ap.parse_args()
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    globals()["MAX_LENGTH"] = args.max_length
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.gpus != world and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29400 + os.getpid() % 500), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

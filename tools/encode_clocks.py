"""SM clock, power and throttle reasons while ONLY the encoder runs back to back for a few seconds (is the tensor phase power-limited?):
python tools/encode_clocks.py [B=64] [seconds=3]"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
from manga_ocr_b200 import crops as C, weights as W
from manga_ocr_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
eng = Engine(W.random_init(0), device=0, max_batch=B, max_length=16)
eng.stage(C.bubble_batch(B)); eng.preprocess(); eng.encode(); eng.sync()
samples, stop = [], False
def poll():
    while not stop:
        try: r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception: r = -1
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, r))
        time.sleep(0.02)
t = threading.Thread(target=poll, daemon=True); t.start()
def phase(name, fn, secs):
    i0 = len(samples); t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < secs:
        for _ in range(20): fn()
        eng.sync(); n += 20
    dt = time.perf_counter() - t0
    s = samples[i0:]
    clk = sorted(x[1] for x in s); pw = [x[2] for x in s]
    reasons = 0
    for x in s: reasons |= max(x[3], 0)
    print(f"{name}: {dt/n*1e3:.3f} ms per call, SM clock median {clk[len(clk)//2]} min {clk[0]} max {clk[-1]} MHz, power mean {sum(pw)/len(pw):.0f} max {max(pw):.0f} W, reasons 0x{reasons:x}", flush=True)
phase("encode x N back to back", eng.encode, secs)
eng.decode(16); eng.sync()
phase("decode(16) back to back", lambda: eng.decode(16), secs)
phase("encode again", eng.encode, secs)
stop = True

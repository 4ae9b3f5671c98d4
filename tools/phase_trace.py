#!/usr/bin/env python3
"""Per-CTA phase timeline of fwd_cols_static (globaltimer stamps): load | barrier | fft | store."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from ka9q_radio_b200 import capi
from ka9q_radio_b200.channelizer import Channelizer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lib = capi.load(); dev = torch.device("cuda:0")
for k, v in [kv.split("=") for kv in sys.argv[2:]]: lib.kgpu_set_tuning(int(k), int(v))
cz = Channelizer(bench.L, bench.M, capi.KGPU_REAL, dev, capacity=4)
host = np.random.default_rng(0).integers(-3000, 3000, 4 * B * bench.L + bench.M - 1, dtype=np.int16)
d_stream = torch.from_numpy(host).to(dev)
spec = cz.alloc_spectra(B)
ncta = 157 * B
dbg = torch.zeros(ncta * 6, dtype=torch.int64, device=dev)
for i in range(3): cz.forward(d_stream, B, spec, scale=bench.SCALE, first_block=i * B)
torch.cuda.synchronize()
ncta2 = 163 * B
dbg2 = torch.zeros(ncta2 * 6, dtype=torch.int64, device=dev)
lib.kgpu_set_debug_buffer(dbg.data_ptr()); lib.kgpu_set_debug_buffer_rows(dbg2.data_ptr())
cz.forward(d_stream, B, spec, scale=bench.SCALE, first_block=0)
torch.cuda.synchronize(); lib.kgpu_set_debug_buffer(None); lib.kgpu_set_debug_buffer_rows(None)
u = dbg2.cpu().numpy().reshape(ncta2, 6).astype(np.float64)
u = u[u[:, 1] > 0]
u0 = u[:, 0].min()
print("ROWS kernel span %.1f us for %d CTAs" % ((u[:, 3].max() - u0) / 1e3, len(u)))
for name, v in (("tma wait", u[:, 1] - u[:, 0]), ("fft(+sync)", u[:, 2] - u[:, 1]), ("epilogue", u[:, 3] - u[:, 2]), ("total", u[:, 3] - u[:, 0])):
    print("  %-13s median %7.2f us  p10 %7.2f  p90 %7.2f" % (name, np.median(v) / 1e3, np.percentile(v, 10) / 1e3, np.percentile(v, 90) / 1e3))
print("COLS")
t = dbg.cpu().numpy().reshape(ncta, 6).astype(np.float64)
t0 = t[:, 0].min()
start, ld, bar, fft, st = t[:, 0] - t0, t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3]
print("kernel span %.1f us for %d CTAs (%d blocks)" % ((t[:, 4].max() - t0) / 1e3, ncta, B))
for name, v in (("load", ld), ("wait@barrier", bar), ("fft", fft), ("store", st), ("total", t[:, 4] - t[:, 0])):
    print("  %-13s median %7.2f us  p10 %7.2f  p90 %7.2f" % (name, np.median(v) / 1e3, np.percentile(v, 10) / 1e3, np.percentile(v, 90) / 1e3))
order = np.argsort(start)
print("  CTA start times (us), every 100th:", np.round(start[order][::100] / 1e3, 1))
sm = t[:, 5].astype(int)
per_sm = np.bincount(sm, minlength=148)
print("  CTAs per SM: min %d max %d" % (per_sm.min(), per_sm.max()))

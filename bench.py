#!/usr/bin/env python3
"""bench.py -- headline benchmark of the overlap-save channelizer hot path (BASELINE.json).

Metric: input Msamples/s through forward + filter + inverse at 1024 NBFM channels (cfg-2:
RX888 129.6 MS/s real int16 input, N = 3 240 000, 1024 x 24 kHz channels, preset fm).
A "step" runs the hot path over `--blocks-per-step` consecutive 20 ms blocks.

  python bench.py --gpus 1 --steps K --warmup W          our arm (CUDA, through the C-ABI)
  python bench.py --impl reference ...                   the reference's own CPU path (oracle/_ref:
                                                         filter.c compiled unmodified + FFT shim)
  torchrun ... bench.py --gpus N                         channel groups sharded over N GPUs, forward
                                                         transform on rank 0 + one NCCL broadcast of
                                                         the spectrum per step (north_star / SURVEY 8e)
One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FS = 129.6e6
L, M = 2592000, 648001          # radio.c:582-587 with blocktime 20 ms, overlap 5
N = L + M - 1
NCHAN = 1024
OLEN = 480                      # 24 kHz * 20 ms
SCALE = float(np.float32(10 ** (3 / 20) / 32768))   # scale_AD(bits=16, real), radio.c:1645-1649
METRIC = "input Msamples/s through forward+filter+inverse at 1024 NBFM channels"


def channel_shift(k: int, group: int = 0) -> int:
    """cfg-2 raster: f_k = 30 MHz + k*25 kHz -> shift 750 000 + 625 k; further
    GPU groups continue the raster (cfg-5 style weak scaling)."""
    return 750_000 + 625 * ((k + NCHAN * group) % 1390)


def make_stream(nblocks: int) -> np.ndarray:
    """Deterministic int16 ADC stream: 16 tones at -30 dBFS on channel centres + noise at -50 dBFS.
    Generated once for 4 blocks by the oracle's sig_gen restatement (xoshiro seed 1) and tiled."""
    from oracle import oracle as O

    base_blocks = min(nblocks, 4)
    f = [(30.0e6 + 25e3 * (64 * i + 3)) / FS for i in range(16)]
    base = O.siggen_tones_i16(base_blocks * L, f, [10 ** (-30 / 20)] * 16, 10 ** (-50 / 20), 1)
    reps = (nblocks + base_blocks - 1) // base_blocks
    return np.tile(base, reps)[: nblocks * L]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
def run_reference(args) -> None:
    """The reference's own CPU implementation of the path, timed on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference_run(blocks=args.ref_blocks, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg-2: RX888 129.6 MS/s real int16->float, N=3240000, 1024 NBFM ch @24 kHz",
                   "blocks_per_step": args.ref_blocks},
        "cpu_baseline": res["cpu_baseline"],
        "e2e": {"value": res["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def cpu_reference_run(blocks: int, steps: int, warmup: int) -> dict:
    """oracle/_ref (reference filter.c unmodified + fftw shim) driven as radiod drives it: one
    producer, FFT worker threads, one pthread per channel (ref_driver.c:ref_bench).  Runs in a
    fresh subprocess because filter.c starts its worker pool once per process (filter.c:1047)."""
    code = r"""
import sys, json, os, time, ctypes as C
sys.path.insert(0, %r)
import numpy as np
from oracle import oracle as O
import bench
R = O.ref_lib()
blocks, steps, warmup, nworkers = %d, %d, %d, %d
x = bench.make_stream(4)
xf, _, _ = O.convert_i16(x, np.float32(bench.SCALE))     # what rx888.c's convert() leaves in the ring
shifts = (C.c_int * bench.NCHAN)(*[bench.channel_shift(k) for k in range(bench.NCHAN)])
drops = C.c_uint(0)
times = []
for s in range(warmup + steps):
    t = R.ref_bench(bench.L, bench.M, O.KO_REAL, bench.NCHAN, bench.OLEN, C.cast(shifts, C.c_void_p), -8000/24000, 8000/24000, 11.0,
                    xf.ctypes.data, len(xf), blocks, nworkers, C.byref(drops))
    if s >= warmup: times.append(t)
print(json.dumps({"times": times, "drops": int(drops.value)}))
"""
    ncores = os.cpu_count() or 1
    nworkers = max(1, min(3, ncores - 1))   # ND=4 spectrum ring: at most 3 forward FFTs in flight (filter.h:48)
    out = subprocess.run([sys.executable, "-c", code % (str(ROOT), blocks, steps, warmup, nworkers)],
                         capture_output=True, text=True, check=True)
    res = json.loads(out.stdout.strip().splitlines()[-1])
    tot = float(np.sum(res["times"]))
    value = blocks * len(res["times"]) * L / tot / 1e6
    from oracle import oracle as O
    kind = "reference" if O.ref_available() else "port"
    return {"value": value, "ms_per_step": 1e3 * tot / len(res["times"]),
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": ncores, "kind": kind,
                             "threads": f"1 producer + {nworkers} forward-FFT workers + {NCHAN} channel threads",
                             "fft_backend": "oracle/fft_cpu.c shim behind fftw3.h (FFTW3 itself is not installed; "
                                            "this is NOT an FFTW-with-wisdom figure)",
                             "sample": f"{blocks} blocks x {len(res['times'])} runs of cfg-2 (float input already in the ring), "
                                       f"{res['drops']} dropped blocks"}}


# --------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from ka9q_radio_b200 import capi
    from ka9q_radio_b200.channelizer import Channelizer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the channelizer has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = capi.load()
    B = args.blocks_per_step
    nstream = args.stream_blocks

    cz = Channelizer(L, M, capi.KGPU_REAL, dev, capacity=NCHAN)
    for k in range(NCHAN):
        cz.add_channel(OLEN, channel_shift(k, rank), -8000 / 24000, 8000 / 24000, 11.0)
    alg_bytes = cz.bank.algorithmic_bytes(capi.KGPU_FMT_I16)       # per block, SURVEY.md 8d
    alg_fwd_in = N * 2.0
    alg_fwd_out = cz.master.bins * 8.0
    alg_chan = alg_bytes - alg_fwd_in - alg_fwd_out

    # ---- inputs: a stream larger than L2, resident in HBM before the timed region ----------
    host = make_stream(nstream)
    hpin = torch.from_numpy(np.concatenate([np.zeros(M - 1, np.int16), host])).pin_memory()
    d_stream = hpin.to(dev)
    ngroups = nstream // B
    from ka9q_radio_b200.sharding import PipelinedSharder

    spec2 = [cz.alloc_spectra(B) for _ in range(2)]
    spec = spec2[0]
    out = cz.alloc_outputs(B)
    comp = torch.cuda.current_stream(dev)
    mc, mc_why = None, "single GPU"
    if world > 1 and args.mg_mode == "spectrum-mc":
        # own multicast copy kernel over NVSwitch when the fabric offers a multicast mapping
        from ka9q_radio_b200.multicast import SpectrumMulticast
        mc, mc_why = SpectrumMulticast.create(rank, world, dev, 2, B * cz.master.spec_stride * 2,
                                                  nctas=int(os.environ.get("KA9Q_MC_CTAS", "64")))
    if mc is not None:
        from ka9q_radio_b200.sharding import MulticastSharder
        symm2 = [mc.slot_view(s, (B, cz.master.spec_stride)) for s in range(2)]
        sharder = MulticastSharder(
            rank, world,
            forward=lambda step, slot: cz.forward(d_stream, B, spec2[slot], scale=SCALE, first_block=(step % ngroups) * B),
            push=lambda slot: mc.push(slot, spec2[slot]),
            ready=mc.ready, arrive=mc.arrive,
            channels=lambda step, slot: cz.channels(spec2[slot] if rank == 0 else symm2[slot], B, out))
    elif args.mg_mode in ("spectrum", "spectrum-mc") or world == 1:
        # north_star: forward transform once (rank 0), ONE broadcast of the block spectra per step
        sharder = PipelinedSharder(
            rank, world,
            forward=lambda step, slot: cz.forward(d_stream, B, spec2[slot], scale=SCALE, first_block=(step % ngroups) * B),
            broadcast=lambda slot: dist.broadcast(spec2[slot], src=0, async_op=True),
            channels=lambda step, slot: cz.channels(spec2[slot], B, out))
    else:
        # alternative: broadcast the raw int16 window (half the bytes) and replicate
        # the forward transform on every GPU
        nwin = (M - 1) + B * L
        win2 = [torch.empty(nwin, dtype=torch.int16, device=dev) for _ in range(2)]

        def stage(step, slot):
            g = (step % ngroups) * B * L
            win2[slot].copy_(d_stream[g:g + nwin])

        def chan_after_forward(step, slot):
            cz.forward(win2[slot], B, spec2[slot], scale=SCALE)
            cz.channels(spec2[slot], B, out)

        sharder = PipelinedSharder(rank, world, forward=stage,
                                   broadcast=lambda slot: dist.broadcast(win2[slot].view(torch.uint8), src=0, async_op=True),
                                   channels=chan_after_forward)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sharder.run(range(args.warmup))
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # hold the same load for ~0.4 s before timing: clocks settle, and nvidia-smi (20 ms period) gets
    # samples under exactly this workload even when the K timed steps last only milliseconds
    hold_ev = torch.cuda.Event(enable_timing=True)
    hold_ev.record(comp)
    nhold = max(args.warmup, int(0.4 / max(1e-4, 2.5e-5 * B)))
    sharder.run(range(nhold))
    barrier()
    lib.kgpu_profile_enable(1)
    lib.kgpu_profile_reset()
    launches0 = lib.kgpu_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(comp)
    sharder.run(range(args.warmup, args.warmup + args.steps))
    e1.record(comp)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.kgpu_launch_count() - launches0
    prof = capi.profile_snapshot()
    lib.kgpu_profile_enable(0)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    samples = args.steps * B * L
    stream_msps = samples / (ms * 1e-3) / 1e6
    value = stream_msps * world          # every rank runs the stream through its own 1024-channel bank

    # ---- end to end through the C-ABI with HOST buffers (pinned), copies inside the timed region
    e2e = run_e2e(args, torch, cz, hpin, dev, rank, world, dist, B, ngroups)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, from the per-launch CUDA-event times ---------------
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback, B200_PROFILING.md)"
    alg_per_kernel = {"fwd_cols": alg_fwd_in * B, "fwd_rows": alg_fwd_out * B, "chan": alg_chan * B}
    kernel_impl = {"fwd_cols": "fwd_cols_v2<int16> (static 1296 = 12.12.9, stage 0 fused with the load, stage 2 with the store)",
                   "fwd_rows": "fwd_rows_v2<real> (static 1250 = 10.25.5, TMA row loads, radix-5 stage fused with the real split)",
                   "chan": "chan_v2<600 = 24.25> (TMA slice+response, product fused into stage 0, output fused into stage 1)"}
    kernels = {}
    for name, (tot_ms, cnt) in prof.items():
        if cnt:
            avg = tot_ms / cnt
            kernels[name] = {"impl": kernel_impl.get(name), "launches": cnt, "avg_ms": avg,
                             "alg_bytes_per_launch": alg_per_kernel.get(name, 0.0),
                             "alg_gbs": alg_per_kernel.get(name, 0.0) / (avg * 1e-3) / 1e9 if avg > 0 else None}
    dom = max(kernels, key=lambda k: kernels[k]["avg_ms"] * kernels[k]["launches"]) if kernels else None
    roof = None
    if dom:
        a = kernels[dom]["alg_gbs"]
        roof = {"kernel": dom, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
                "traffic": (TRAFFIC_NCU.get(dom) * B) if isinstance(TRAFFIC_NCU.get(dom), (int, float)) else None,
                "traffic_note": "ncu dram__bytes_read+write per block (profiles/traffic.json) x blocks per launch",
                "peak_source": peak_src,
                "note": "algorithmic bytes attributed per kernel: fwd_cols = window read (N*2 B), fwd_rows = spectrum "
                        "write (bins*8 B), chan = slices+responses+outputs; the inter-pass buffer earns no credit"}
    sum_ms = sum(v["avg_ms"] for k, v in kernels.items() if k in alg_per_kernel)
    pipeline = {"alg_bytes_per_step": alg_bytes * B, "kernel_ms_per_step": sum_ms,
                "achieved": alg_bytes * B / (sum_ms * 1e-3) / 1e9 if sum_ms else None, "unit": "GB/s"}
    if pipeline["achieved"]:
        pipeline["frac"] = pipeline["achieved"] / peak
    pipeline["achieved_wall"] = alg_bytes * B * args.steps / (ms * 1e-3) / 1e9
    pipeline["frac_wall"] = pipeline["achieved_wall"] / peak

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_run(blocks=args.ref_blocks, steps=2, warmup=1)["cpu_baseline"]
        except Exception as ex:  # the GPU numbers stand on their own
            cpu = {"value": None, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}

    line = {
        "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg-2: RX888 129.6 MS/s real int16, N=3240000 (L=2592000, M=648001), "
                               f"{NCHAN} NBFM ch @24 kHz (Ns=600) per GPU",
                   "blocks_per_step": B, "input_stream_MB": round(host.nbytes / 1e6, 1),
                   "l2_policy": "input stream larger than L2 (126 MB), consecutive groups cycled; no explicit flush",
                   "plan": cz.master.describe(),
                   "parallelism": ("single GPU" if world == 1 else
                                   (f"{world} GPUs: forward on rank 0, spectra stored once per step to an NVSwitch multicast "
                                    "address by kgpu_multicast_copy (own kernel, multimem.st) + symmetric-memory barrier, "
                                    if mc is not None else
                                    f"{world} GPUs: forward on rank 0, 1 NCCL broadcast of the spectrum per step"
                                    + (f" (multicast path unavailable: {mc_why}), " if args.mg_mode == "spectrum-mc" else ", ")
                                    if args.mg_mode in ("spectrum", "spectrum-mc") else
                                    f"{world} GPUs: 1 NCCL broadcast of the int16 window per step, forward replicated, ")
                                   + f"{NCHAN} channels per GPU; value = stream rate x GPUs"),
                   "stream_msps": stream_msps, "realtime_factor": stream_msps / 129.6},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roof,
        "roofline_pipeline": pipeline, "kernels": kernels, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per launch (filled from profiles/ once captured)
TRAFFIC_NCU: dict = {}
try:
    TRAFFIC_NCU = json.loads((ROOT / "profiles" / "traffic.json").read_text())
except Exception:
    pass


def run_e2e(args, torch, cz, hpin, dev, rank, world, dist, B, ngroups) -> dict:
    """Same metric through the C-ABI with HOST buffers: every step copies that step's window
    (M-1 history + B*L new int16 samples) from pinned host memory, runs forward + channels and
    reads the channel outputs back into pinned host memory.  Copies are double-buffered on
    separate streams (a streaming receiver would do the same); all of it is inside the timed region."""
    from ka9q_radio_b200 import capi

    nwin = (M - 1) + B * L
    d_in = [torch.empty(nwin, dtype=torch.int16, device=dev) for _ in range(2)]
    d_out = [cz.alloc_outputs(B) for _ in range(2)]
    d_spec = [cz.alloc_spectra(B) for _ in range(2)]
    h_out = [torch.empty(d_out[0].shape, dtype=torch.complex64).pin_memory() for _ in range(2)]
    s_in, s_comp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_comp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    steps, warm = args.steps, args.warmup

    def one(s):
        j = s & 1
        g = s % ngroups
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_comp[j])               # previous use of d_in[j] finished
            d_in[j].copy_(hpin[g * B * L: g * B * L + nwin], non_blocking=True)
            ev_in[j].record(s_in)
        with torch.cuda.stream(s_comp):
            s_comp.wait_event(ev_in[j])
            s_comp.wait_event(ev_out[j])              # previous D2H of d_out[j] finished
            if rank == 0 or world == 1:
                cz.forward(d_in[j], B, d_spec[j], scale=SCALE)
            if world > 1:
                dist.broadcast(d_spec[j], src=0)
            cz.channels(d_spec[j], B, d_out[j])
            ev_comp[j].record(s_comp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_comp[j])
            h_out[j].copy_(d_out[j], non_blocking=True)
            ev_out[j].record(s_out)

    for s in range(warm):
        one(s)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(s_in)
    for s in range(steps):
        one(warm + s)
    s_out.synchronize()
    s_comp.synchronize()
    t1.record(s_out)
    torch.cuda.synchronize(dev)
    ms = t0.elapsed_time(t1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    v = steps * B * L / (ms * 1e-3) / 1e6 * world
    return {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": int(nwin * 2),
            "d2h_bytes_per_step": int(d_out[0].numel() * 8), "ms_per_step": ms / steps,
            "api": "kgpu_forward + kgpu_bank_run (C-ABI, include/ka9q_gpu.h) on pinned host buffers, "
                   "H2D/compute/D2H double-buffered on 3 streams"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--blocks-per-step", type=int, default=32)
    ap.add_argument("--stream-blocks", type=int, default=64, help="resident input stream length (blocks)")
    ap.add_argument("--ref-blocks", type=int, default=12, help="blocks per step of the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mg-mode", default="spectrum", choices=["spectrum", "spectrum-mc", "input"],
                    help="multi-GPU hand-off: NCCL broadcast of the forward spectrum (north_star, default); the same through this "
                         "repository's NVSwitch-multicast copy kernel; or NCCL broadcast of the raw input window with the "
                         "forward transform replicated")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

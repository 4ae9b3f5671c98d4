"""BASELINE.json's configurations as data: geometry, channel plan and a deterministic synthetic input stream.

Host-side only (numpy), no checker code.  SURVEY.md section 8d defines the inputs:
  cfg1  sig_gen real 2.4 MS/s, 1 NBFM channel at 600 kHz
  cfg2  RX888 129.6 MS/s real int16, 1024 NBFM channels f_k = 30 MHz + k*25 kHz (+ optionally 8 inverted ones)
  cfg3  RX888, 300 SSB channels, 100 each at 12 / 24 / 48 kHz, preset usb, 90 kHz raster from 1.8 MHz
  cfg4  20 MS/s complex int16 I/Q, 512 NBFM channels on a 25 kHz raster across -6.4 .. +6.4 MHz
  cfg5  RX888, 8192 NBFM channels f_k = 0.5 MHz + k*7.52 kHz (188 bins), preset nfm, 1024 contiguous channels per GPU
Presets: share/presets.conf:67-82 (fm: 24 kHz, -8k..+8k), :84-89 (nfm: 24 kHz, +-6.25 kHz), :236-242 (usb: +50..+3000 Hz);
Kaiser beta 11 (modes.c:40).  Block time 20 ms, overlap 5 (radio.c:59,71): L = fs/50, M = L/4 + 1.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

KGPU_COMPLEX, KGPU_REAL = 1, 2


@dataclass
class Channel:
    olen: int
    shift: int
    low: float     # fractions of the output rate (set_filter's convention, filter.c:968)
    high: float
    beta: float = 11.0


@dataclass
class Workload:
    name: str
    description: str
    fs: float
    in_type: int
    L: int
    M: int
    scale: float
    channels: list = field(default_factory=list)
    tones_hz: list = field(default_factory=list)

    @property
    def N(self) -> int:
        return self.L + self.M - 1

    @property
    def samples_per_block(self) -> int:   # int16 words per block (I/Q pairs count twice)
        return self.L * (2 if self.in_type == KGPU_COMPLEX else 1)

    def shift_of(self, f_hz: float) -> int:
        """compute_tuning (radio.c:1175-1199): nearest bin"""
        return int(np.rint(f_hz / (self.fs / self.N)))

    def stream(self, nblocks: int, seed: int = 1) -> np.ndarray:
        """int16 ADC words for `nblocks` blocks: tones at -30 dBFS + Gaussian noise at -50 dBFS, built for 4 blocks and tiled
        (the tones sit on exact multiples of 50/4 Hz... not needed: tiling only has to be deterministic, not continuous)."""
        base = min(nblocks, 4)
        n = base * self.L
        rng = np.random.default_rng(seed)
        t = np.arange(n, dtype=np.float64)
        amp, sigma = 10 ** (-30 / 20), 10 ** (-50 / 20)
        if self.in_type == KGPU_REAL:
            x = sigma * rng.standard_normal(n)
            for k, f in enumerate(self.tones_hz):
                x += amp * np.cos(2 * np.pi * ((f / self.fs * t) % 1.0) + 0.7 * k)
            w = np.clip(np.rint(32767.0 * x), -32767, 32767).astype(np.int16)
        else:
            z = sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2)
            for k, f in enumerate(self.tones_hz):
                z += amp * np.exp(2j * np.pi * ((f / self.fs * t) % 1.0) + 0.7j * k)
            w = np.empty(2 * n, np.int16)
            w[0::2] = np.clip(np.rint(32767.0 * z.real), -32767, 32767)
            w[1::2] = np.clip(np.rint(32767.0 * z.imag), -32767, 32767)
        reps = (nblocks + base - 1) // base
        return np.tile(w, reps)[: nblocks * self.samples_per_block]


def _geometry(fs: float) -> tuple[int, int]:
    L = int(round(fs * 0.020))   # radio.c:582
    return L, L // 4 + 1          # overlap 5: M - 1 = L / 4 (radio.c:583-586)


_SCALE_REAL16 = float(np.float32(10 ** (3 / 20) / 32768))   # scale_AD(bits=16, real), radio.c:1645-1649
_SCALE_CPLX16 = float(np.float32(1.0 / 32768))


def cfg2(group: int = 0, with_inverted: bool = False) -> Workload:
    fs = 129.6e6
    L, M = _geometry(fs)
    w = Workload("cfg2", "cfg-2: RX888 129.6 MS/s real int16, N=3240000 (L=2592000, M=648001), 1024 NBFM ch @24 kHz (Ns=600)",
                 fs, KGPU_REAL, L, M, _SCALE_REAL16)
    for k in range(1024):
        w.channels.append(Channel(480, 750_000 + 625 * ((k + 1024 * group) % 1390), -8000 / 24000, 8000 / 24000))
    if with_inverted:  # what a high-side-injection tuner produces (filter.c:814, :856-892)
        for k in (0, 3, 64, 100, 511, 700, 900, 1023):
            w.channels.append(Channel(480, -(750_000 + 625 * k), -8000 / 24000, 8000 / 24000))
    w.tones_hz = [30.0e6 + 25e3 * (64 * i + 3) for i in range(16)]
    return w


def cfg3() -> Workload:
    fs = 129.6e6
    L, M = _geometry(fs)
    w = Workload("cfg3", "cfg-3: RX888 129.6 MS/s real int16, 300 SSB ch (100 each @12/24/48 kHz, preset usb), 90 kHz raster from 1.8 MHz",
                 fs, KGPU_REAL, L, M, _SCALE_REAL16)
    for i in range(300):
        rate = (12000, 24000, 48000)[i % 3]
        w.channels.append(Channel(rate // 50, w.shift_of(1.8e6 + 90e3 * i), 50.0 / rate, 3000.0 / rate))
    w.tones_hz = [1.8e6 + 90e3 * (19 * i + 2) + 1000.0 for i in range(15)]
    return w


def cfg4() -> Workload:
    fs = 20e6
    L, M = _geometry(fs)
    w = Workload("cfg4", "cfg-4: 20 MS/s complex int16 I/Q, N=500000 c2c (L=400000, M=100001), 512 NBFM ch @24 kHz, 25 kHz raster -6.4..+6.4 MHz",
                 fs, KGPU_COMPLEX, L, M, _SCALE_CPLX16)
    for k in range(512):
        w.channels.append(Channel(480, w.shift_of(-6.4e6 + 25e3 * k), -8000 / 24000, 8000 / 24000))
    w.tones_hz = [-6.4e6, -1.0e6 + 25e3, 25e3 * 7, 3.2e6, 6.375e6]
    return w


def cfg5(rank: int = 0, world: int = 8) -> Workload:
    """one GPU's share of the 8192-channel plan: channels 1024*rank .. 1024*rank + 1023"""
    fs = 129.6e6
    L, M = _geometry(fs)
    w = Workload("cfg5", f"cfg-5: RX888 129.6 MS/s real int16, 8192 NBFM ch f_k = 0.5 MHz + k*7.52 kHz (preset nfm), "
                         f"ch {1024 * rank}..{1024 * rank + 1023} on this GPU (of {world})",
                 fs, KGPU_REAL, L, M, _SCALE_REAL16)
    for k in range(1024 * rank, 1024 * (rank + 1)):
        w.channels.append(Channel(480, w.shift_of(0.5e6 + 7520.0 * k), -6250 / 24000, 6250 / 24000))
    w.tones_hz = [0.5e6 + 7520.0 * (512 * i + 5) for i in range(16)]
    return w


def by_name(name: str, rank: int = 0, world: int = 1) -> Workload:
    if name == "cfg2":
        return cfg2(rank)
    if name == "cfg3":
        return cfg3()
    if name == "cfg4":
        return cfg4()
    if name == "cfg5":
        return cfg5(rank, max(world, 1))
    raise ValueError(f"unknown workload {name}")

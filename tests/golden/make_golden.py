#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the REFERENCE's own filter.c (oracle/_ref, compiled
unmodified from /root/reference/src by oracle/Makefile) -- the reference ships no golden vectors
or tests of its own, so these fixtures are its outputs captured here.

Run in the dev container (needs /root/reference):  python tests/golden/make_golden.py
The fixtures are small on purpose; the input streams are regenerated from the stored sig_gen
parameters (the restated generator is bit-exact with sig_gen.c/gauss.c/osc.c, which the fixture
also pins via a few raw samples and a float64 checksum).
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent

CASES = {
    # name: (in_type, L, M, nblocks, siggen(amplitude, noise, cycles/sample, scale), notch bins, channels)
    "real_small": dict(
        in_type=O.KO_REAL, L=4800, M=1201, nb=4, sig=(0.1, 0.01, 0.25 + 1 / 64.0, 10 ** (3 / 20)), notch=[77],
        chans=[dict(olen=48, shift=1594, low=-1 / 3, high=1 / 3, beta=11.0),
               dict(olen=48, shift=-1594, low=-1 / 3, high=1 / 3, beta=11.0),
               dict(olen=96, shift=1590, low=0.01, high=0.3, beta=7.0),
               dict(olen=24, shift=3, low=-0.4, high=0.4, beta=3.0),
               dict(olen=48, shift=1600, low=-0.25, high=0.25, beta=11.0, isb=True)]),
    "complex_small": dict(
        in_type=O.KO_COMPLEX, L=4000, M=1001, nb=4, sig=(0.1, 0.01, -0.123, 1.0), notch=None,
        chans=[dict(olen=80, shift=-615, low=-1 / 3, high=1 / 3, beta=11.0),
               dict(olen=80, shift=2499, low=-1 / 3, high=1 / 3, beta=11.0),
               dict(olen=160, shift=-2490, low=-0.2, high=0.45, beta=5.0),
               dict(olen=40, shift=0, low=-0.5, high=0.5, beta=11.0)]),
    # cfg-1 of BASELINE.json: sig_gen real 2.4 MS/s, carrier 600 kHz at -20 dBFS, noise -40 dBFS,
    # one 24 kHz NBFM channel (preset fm: +-8 kHz, beta 11), DC notch as radio.c:601-620 installs it
    "cfg1_siggen": dict(
        in_type=O.KO_REAL, L=48000, M=12001, nb=6, sig=(10 ** (-20 / 20), 10 ** (-40 / 20), 600e3 / 2.4e6, 10 ** (3 / 20)),
        notch=[], chans=[dict(olen=480, shift=15000, low=-8000 / 24000, high=8000 / 24000, beta=11.0)]),
}


def make_stream(c):
    a, n, f, s = c["sig"]
    if c["in_type"] == O.KO_REAL:
        return O.ref_siggen_real(c["nb"] * c["L"], a, n, f, s)
    return O.ref_siggen_complex(c["nb"] * c["L"], a, n, f, s)


def main():
    for name, c in CASES.items():
        x = make_stream(c)
        outs, specs = O.ref_run_stream(x, c["L"], c["M"], c["chans"], notch_bins=c["notch"], keep_spectra=True)
        with O.RefSession(c["L"], c["M"], c["in_type"]) as s:
            resp = [s.response(s.add_channel(ch["olen"], ch["low"], ch["high"], ch["beta"])) for ch in c["chans"]]
        data = dict(
            in_type=c["in_type"], L=c["L"], M=c["M"], nb=c["nb"], sig=np.array(c["sig"], np.float64),
            notch=np.array(c["notch"] if c["notch"] is not None else [-1], np.int64),
            chan_params=np.array([[ch["olen"], ch["shift"], ch["low"], ch["high"], ch["beta"], float(ch.get("isb", False))]
                                  for ch in c["chans"]], np.float64),
            x_head=x[:16].copy(), x_sum=np.array([np.sum(x.astype(np.complex128))]),
            x_sha=np.frombuffer(hashlib.sha256(x.tobytes()).digest(), np.uint8),
        )
        stride = max(1, len(specs[0]) // 512)
        data["spec_stride"] = stride
        data["spec_sub"] = np.stack([sp[::stride] for sp in specs])
        data["spec_absmax"] = np.array([np.abs(sp).max() for sp in specs])
        data["spec_sum"] = np.array([np.sum(sp.astype(np.complex128)) for sp in specs])
        data["spec_energy"] = np.array([np.sum(np.abs(sp.astype(np.complex128)) ** 2) for sp in specs])
        for i in range(len(c["chans"])):
            data[f"out{i}"] = np.stack([outs[b][i] for b in range(c["nb"])])
            data[f"resp{i}"] = resp[i]
        np.savez_compressed(OUT / f"{name}.npz", **data)
        print(name, "written", sum(v.nbytes for v in data.values() if hasattr(v, "nbytes")) // 1024, "KiB raw")


if __name__ == "__main__":
    main()

"""Torch-tensor convenience wrapper over the C-ABI: device memory, streams and nothing else.

Mirrors how radiod drives filter.h (reference radio.c:582-620 setup, rx888.c:800-826 producer,
radio.c:1460 consumer) but batched: one forward launch pair and one channel launch per group of
blocks.  All arithmetic happens in libka9qgpu.so; torch only owns the buffers.
"""
from __future__ import annotations

import numpy as np
import torch

from . import capi


class Channelizer:
    def __init__(self, L: int, M: int, in_type: int, device: str | torch.device = "cuda:0", capacity: int = 1024):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise capi.KgpuError("Channelizer needs a CUDA device; there is no CPU fallback")
        torch.cuda.set_device(self.device)
        capi.check(capi.load().kgpu_set_device(self.device.index or 0), "kgpu_set_device")
        self.master = capi.Master(L, M, in_type)
        self.bank = capi.Bank(self.master, capacity)
        self.L, self.M, self.N, self.in_type = L, M, L + M - 1, in_type
        self.nchan = 0
        self._olen = {}

    # ---- channel management (create_filter_output + set_filter + shift) -------------------
    def add_channel(self, olen, shift, low=None, high=None, beta=None, response=None, isb=False) -> int:
        idx = self.nchan
        pts = self.bank.define(idx, olen)
        if response is not None:
            self.bank.set_response(idx, response)
        else:
            self.bank.set_filter(idx, low, high, beta)
        self.bank.set_shift(idx, shift)
        if isb:
            self.bank.set_flags(idx, capi.KGPU_CHAN_ISB)
        self._olen[idx] = (olen, pts)
        self.nchan += 1
        return idx

    # ---- data movement helpers -------------------------------------------------------------
    def stage_stream(self, samples: np.ndarray) -> torch.Tensor:
        """Host stream of nblocks*L new samples -> device tensor with the M-1 zero history the
        reference's zeroed ring provides at start-up (filter.c:242-244, :257-259)."""
        npad = self.M - 1
        if self.in_type == capi.KGPU_COMPLEX and samples.dtype == np.int16:
            npad *= 2  # interleaved I/Q
        pad = np.zeros(npad, samples.dtype)
        t = torch.from_numpy(np.concatenate([pad, samples]))
        return t.to(self.device)

    def fmt_of(self, t: torch.Tensor) -> int:
        return capi.KGPU_FMT_I16 if t.dtype == torch.int16 else capi.KGPU_FMT_F32

    def alloc_spectra(self, nblocks) -> torch.Tensor:
        return torch.empty((nblocks, self.master.spec_stride), dtype=torch.complex64, device=self.device)

    def alloc_outputs(self, nblocks) -> torch.Tensor:
        return torch.empty((nblocks, max(self.bank.out_stride, 1)), dtype=torch.complex64, device=self.device)

    # ---- the two halves of the path ----------------------------------------------------------
    def forward(self, d_stream: torch.Tensor, nblocks: int, spectra: torch.Tensor, scale: float = 1.0,
                first_block: int = 0, derandomize: bool = False, stats: torch.Tensor | None = None) -> None:
        i16 = d_stream.dtype == torch.int16
        if self.in_type == capi.KGPU_COMPLEX:
            esz = 4 if i16 else 8   # one I/Q pair
        else:
            esz = 2 if i16 else 4
        off = first_block * self.L * esz
        st = torch.cuda.current_stream(self.device).cuda_stream
        self.master.forward(d_stream.data_ptr() + off, self.fmt_of(d_stream), scale, nblocks, spectra.data_ptr(), st,
                            derandomize, stats.data_ptr() if stats is not None else 0)

    def apply_notches(self, spectra: torch.Tensor, nblocks: int) -> None:
        st = torch.cuda.current_stream(self.device).cuda_stream
        self.master.apply_notches(spectra.data_ptr(), nblocks, st)

    def channels(self, spectra: torch.Tensor, nblocks: int, outputs: torch.Tensor) -> None:
        st = torch.cuda.current_stream(self.device).cuda_stream
        self.bank.run(spectra.data_ptr(), nblocks, outputs.data_ptr(), st)

    def channel_slice(self, outputs: torch.Tensor, idx: int) -> torch.Tensor:
        off = self.bank.out_offset(idx)
        return outputs[:, off:off + self._olen[idx][0]]

    def close(self):
        self.bank.close()
        self.master.close()

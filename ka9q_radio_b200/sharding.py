"""Multi-GPU orchestration of the channelizer: shard independent channel groups over ranks.

The path partitions naturally: every slave only reads the master's spectrum
(reference filter.c:703-707), so channels are split into contiguous groups of ceil(C/G) per GPU,
rank 0 runs the forward transform and ONE broadcast of the spectrum per step hands it to the
other ranks (north_star).  This module is pure host logic -- the compute and the collective are
injected -- so the same code runs under NCCL on GPUs (bench.py) and under gloo on CPUs (tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Sequence


def channel_groups(nchan: int, world: int) -> list[range]:
    """Contiguous groups of ceil(nchan/world) channels; trailing ranks may get fewer (or none)."""
    if world < 1 or nchan < 0:
        raise ValueError("bad partition request")
    per = -(-nchan // world) if nchan else 0
    return [range(min(r * per, nchan), min((r + 1) * per, nchan)) for r in range(world)]


@dataclass
class PipelinedSharder:
    """Two-deep pipeline: while rank 0 transforms step s and the broadcast of s is in flight, every
    rank runs its channel group on the spectrum of step s-1.

    forward(step, slot)         rank 0 only: fill spectrum buffer `slot` with step's block spectra
    broadcast(slot) -> handle   collective from rank 0 over buffer `slot`; returns an object with .wait()
                                (torch.distributed Work) or None when synchronous
    channels(step, slot)        this rank's channel group on buffer `slot`
    """

    rank: int
    world: int
    forward: Callable[[int, int], None]
    broadcast: Callable[[int], object]
    channels: Callable[[int, int], None]
    depth: int = 2
    forward_on_all: bool = False   # block-parallel forward: every rank transforms its share, the collective is an all-gather

    def run(self, steps: Sequence[int]) -> None:
        pending: list[tuple[int, int, object]] = []
        for i, step in enumerate(steps):
            slot = i % self.depth
            if self.rank == 0 or self.forward_on_all:
                self.forward(step, slot)
            handle = self.broadcast(slot) if self.world > 1 else None
            pending.append((step, slot, handle))
            if len(pending) >= self.depth:
                self._drain_one(pending)
        while pending:
            self._drain_one(pending)

    def _drain_one(self, pending) -> None:
        step, slot, handle = pending.pop(0)
        if handle is not None:
            handle.wait()
        self.channels(step, slot)


@dataclass
class SliceExchange:
    """Block-parallel forward + slice hand-off (bench.py --mg-mode a2a): every rank transforms `blocks_per_rank` of a step's
    blocks and ONE all-to-all sends each peer only the bins that peer's channels read.

    windows[r] = (lo, hi): the bin range rank r's channels read (slave walk, reference filter.c:728-893, plus the half
    transform on either side).  send / recv buffers are (world, blocks_per_rank, width) complex tensors; a step's spectra are
    rows p*blocks_per_rank + b of a (world*blocks_per_rank, spec_stride) tensor.  When the windows are regularly spaced (equal
    contiguous channel groups) the pack and the unpack are one strided copy each instead of `world` copies: at 4-8 GPUs the
    step is paced by the host issuing launches, not by the GPUs."""

    rank: int
    world: int
    windows: Sequence[tuple[int, int]]
    spec_stride: int
    blocks_per_rank: int

    def __post_init__(self):
        self.width = max(hi - lo for lo, hi in self.windows)
        los = [lo for lo, _ in self.windows]
        self.gap = los[1] - los[0] if self.world > 1 else 0
        self.regular = (self.world > 1 and self.gap > 0 and all(los[q] == los[0] + q * self.gap for q in range(self.world))
                        and los[-1] + self.width <= self.spec_stride)

    def pack(self, mine, send) -> None:
        """mine: this rank's (blocks_per_rank, spec_stride) spectra -> send[q, b, :] = mine[b, lo_q : lo_q + width]"""
        import torch

        if self.regular:
            send.copy_(torch.as_strided(mine, (self.world, self.blocks_per_rank, self.width), (self.gap, self.spec_stride, 1),
                                        mine.storage_offset() + self.windows[0][0]))
            return
        for q, (lo, hi) in enumerate(self.windows):
            send[q, :, : hi - lo].copy_(mine[:, lo:hi])

    def unpack(self, spectra, recv) -> None:
        """recv[p, b, :] (from rank p) -> spectra[p*blocks_per_rank + b, lo : hi] for this rank's window"""
        import torch

        lo, hi = self.windows[self.rank]
        if self.regular:
            torch.as_strided(spectra, (self.world, self.blocks_per_rank, self.width),
                             (self.blocks_per_rank * self.spec_stride, self.spec_stride, 1), spectra.storage_offset() + lo).copy_(recv)
            return
        n = self.blocks_per_rank
        for p in range(self.world):
            spectra[p * n:(p + 1) * n, lo:hi].copy_(recv[p, :, : hi - lo])


@dataclass
class MulticastSharder:
    """Spectrum hand-off through an NVSwitch multicast mapping instead of a collective call.

    rank 0 keeps its spectra in a private buffer, and after each forward transform its own copy
    kernel (kgpu_multicast_copy, on a side stream) stores them once to a multicast address that
    lands in the symmetric buffer of every GPU, followed by the group barrier; the copy overlaps the
    forward transform of the next step.  The peers enter the same barrier on their compute stream
    and then run their channel group on their local copy.

    forward(step, slot)    rank 0: fill the private spectrum buffer `slot`
    push(slot)             rank 0: side stream: wait for forward, copy to multicast, group barrier
    ready(slot)            rank 0: make the compute stream wait for push(slot) (buffer reusable)
    arrive(slot)           peers: the group barrier on the compute stream
    channels(step, slot)   this rank's channel group (rank 0 reads its private buffer)

    Two slots are enough: push(k+2) is ordered behind barrier k+1, which a peer only enters after
    its channels(k) -- the last reader of the slot that push(k+2) overwrites.
    """

    rank: int
    world: int
    forward: Callable[[int, int], None]
    push: Callable[[int], None]
    ready: Callable[[int], None]
    arrive: Callable[[int], None]
    channels: Callable[[int, int], None]

    def run(self, steps: Sequence[int]) -> None:
        prev = None
        for i, step in enumerate(steps):
            slot = i % 2
            if self.rank != 0:
                self.arrive(slot)
                self.channels(step, slot)
                continue
            self.forward(step, slot)
            if self.world > 1:
                self.push(slot)
            if prev is not None:
                if self.world > 1:
                    self.ready(prev[1])
                self.channels(*prev)
            prev = (step, slot)
        if self.rank == 0 and prev is not None:
            if self.world > 1:
                self.ready(prev[1])
            self.channels(*prev)

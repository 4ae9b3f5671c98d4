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

    def run(self, steps: Sequence[int]) -> None:
        pending: list[tuple[int, int, object]] = []
        for i, step in enumerate(steps):
            slot = i % self.depth
            if self.rank == 0:
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

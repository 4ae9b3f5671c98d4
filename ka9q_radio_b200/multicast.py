"""Spectrum hand-off between GPUs over an NVSwitch multicast mapping.

torch only does the plumbing here: it allocates the symmetric buffer, exchanges the handles
(torch.distributed._symmetric_memory) and provides the group barrier; the bytes are moved by this
repository's own kernel (kgpu_multicast_copy -> mc_push_kernel, multimem.st): rank 0 stores every
spectrum ONCE and the switch replicates it into the symmetric buffer of every GPU.

`SpectrumMulticast.create()` returns None when the fabric, the driver or torch cannot provide a
multicast mapping (single GPU, no NVSwitch, no symmetric-memory support); the caller then keeps the
NCCL broadcast.  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import capi


class SpectrumMulticast:
    def __init__(self, rank, world, device, symm, hdl, mc_base, nctas):
        self.rank, self.world, self.device = rank, world, device
        self.symm = symm            # float32 [slots, floats_per_slot], symmetric across the group
        self.hdl = hdl
        self.mc_base = mc_base      # multicast address of symm[0, 0]
        self.nctas = nctas
        self.lib = capi.load()
        self.side = torch.cuda.Stream(device=device, priority=-1) if rank == 0 else None
        self.done = [torch.cuda.Event() for _ in range(symm.shape[0])]
        self.slot_bytes = symm.shape[1] * 4

    # ---- construction ---------------------------------------------------------------------------
    @classmethod
    def create(cls, rank: int, world: int, device, slots: int, floats_per_slot: int, group=None, nctas: int = 64):
        """Collective over `group` (default: WORLD).  Returns (object, "ok"), or (None, reason) on every
        rank when a multicast mapping cannot be had."""
        import torch.distributed as dist

        ok, obj, why = 1, None, ""
        try:
            import torch.distributed._symmetric_memory as symm_mem

            group = group if group is not None else dist.group.WORLD
            floats_per_slot = (floats_per_slot + 3) // 4 * 4
            symm = symm_mem.empty((slots, floats_per_slot), dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(symm, group)
            mc = int(hdl.multicast_ptr or 0)
            if mc == 0:
                raise RuntimeError("no multicast mapping (multicast_ptr == 0)")
            # the handle describes the whole allocation; this tensor may start `off` bytes into it
            off = int(symm.data_ptr()) - int(hdl.buffer_ptrs[rank])
            obj = cls(rank, world, device, symm, hdl, mc + off, nctas)
        except Exception as ex:  # noqa: BLE001 - any failure means "use NCCL"
            ok, why = 0, f"{type(ex).__name__}: {ex}"
        t = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) == 0:
            return None, (why or "unavailable on a peer")
        if not obj._self_test():
            return None, "multicast self-test failed"
        return obj, "ok"

    def _self_test(self) -> bool:
        """Rank 0 stores a pattern through the multicast address; every rank must see it locally."""
        import torch.distributed as dist

        n = 4096
        self.symm.zero_()
        torch.cuda.synchronize(self.device)
        self.hdl.barrier(channel=0)
        if self.rank == 0:
            pat = torch.arange(n, dtype=torch.float32, device=self.device) + 1.0
            last = self.symm.shape[0] - 1
            capi.check(self.lib.kgpu_multicast_copy(pat.data_ptr(), self.mc_base + last * self.slot_bytes, n * 4, 4,
                                                    torch.cuda.current_stream(self.device).cuda_stream), "kgpu_multicast_copy")
        self.hdl.barrier(channel=0)
        torch.cuda.synchronize(self.device)
        want = torch.arange(n, dtype=torch.float32, device=self.device) + 1.0
        good = bool(torch.equal(self.symm[-1, :n], want)) and float(self.symm[-1, n:2 * n].abs().max()) == 0.0
        t = torch.tensor([1 if good else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        self.symm.zero_()
        torch.cuda.synchronize(self.device)
        self.hdl.barrier(channel=0)
        return int(t.item()) == 1

    # ---- per-step operations (MulticastSharder callbacks) -------------------------------------------
    def slot_view(self, slot: int, shape, dtype=torch.complex64) -> torch.Tensor:
        """This rank's symmetric copy of `slot`, viewed as the spectra tensor the channel bank reads."""
        n = 1
        for d in shape:
            n *= d
        flat = self.symm[slot, : n * 2]
        return torch.view_as_complex(flat.view(*shape, 2)) if dtype == torch.complex64 else flat.view(*shape)

    def push(self, slot: int, src: torch.Tensor) -> None:
        """rank 0: after the work queued on the current stream, copy `src` to every GPU's slot and signal."""
        nbytes = src.numel() * src.element_size()
        assert nbytes <= self.slot_bytes and nbytes % 16 == 0
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            capi.check(self.lib.kgpu_multicast_copy(src.data_ptr(), self.mc_base + slot * self.slot_bytes, nbytes, self.nctas,
                                                    self.side.cuda_stream), "kgpu_multicast_copy")
            self.hdl.barrier(channel=0)
            self.done[slot].record(self.side)

    def ready(self, slot: int) -> None:
        torch.cuda.current_stream(self.device).wait_event(self.done[slot])

    def arrive(self, slot: int) -> None:
        self.hdl.barrier(channel=0)

"""N>1 host logic on CPU: world_size-2 gloo run of the channel-sharding pipeline.

The compute callbacks here are the ORACLE (test infrastructure); what is under test is the
orchestration in ka9q_radio_b200/sharding.py that bench.py uses with the CUDA kernels + NCCL:
partitioning, the two-deep forward/broadcast/channels pipeline, buffer-slot reuse."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_channel_groups():
    from ka9q_radio_b200.sharding import channel_groups

    g = channel_groups(8192, 8)
    assert [len(r) for r in g] == [1024] * 8 and g[3][0] == 3072
    g = channel_groups(10, 4)
    assert [list(r) for r in g] == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9]]
    assert [len(r) for r in channel_groups(2, 4)] == [1, 1, 0, 0]
    assert [len(r) for r in channel_groups(0, 2)] == [0, 0]


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from ka9q_radio_b200.sharding import PipelinedSharder, channel_groups
    from oracle import oracle as O

    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, M, nsteps = 4800, 1201, 5
    N = L + M - 1
    x = O.siggen_real(nsteps * L, 0.1, 0.02, 0.27, 1.0)
    chans = [dict(olen=48, shift=900 + 200 * i, low=-0.3, high=0.3, beta=9.0) for i in range(7)]
    mine = channel_groups(len(chans), world)[rank]
    resp = {i: O.design_response(60, 48, N, True, -0.3, 0.3, 9.0) for i in mine}
    spec = [torch.zeros(N // 2 + 1, dtype=torch.complex64) for _ in range(2)]
    out = {}

    def forward(step, slot):
        spec[slot].copy_(torch.from_numpy(O.forward(O.block_window(x, L, M, step))))

    def broadcast(slot):
        return dist.broadcast(spec[slot], src=0, async_op=True)

    def channels(step, slot):
        X = spec[slot].numpy()
        for i in mine:
            out[(step, i)] = O.channel_block(O.KO_REAL, X, resp[i], chans[i]["shift"])[-48:].copy()

    if os.environ.get("KA_SHARDER") == "multicast":
        # the NVSwitch multicast store + group barrier of bench.py, emulated by a blocking gloo
        # broadcast: rank 0 keeps private spectra (spec), the peers read the "symmetric" copy (symm)
        from ka9q_radio_b200.sharding import MulticastSharder

        symm = [torch.zeros(N // 2 + 1, dtype=torch.complex64) for _ in range(2)]
        log = []

        def push(slot):
            log.append(("push", slot))
            symm[slot].copy_(spec[slot])
            dist.broadcast(symm[slot], src=0)

        def arrive(slot):
            dist.broadcast(symm[slot], src=0)

        def channels_mc(step, slot):
            X = (spec if rank == 0 else symm)[slot].numpy()
            for i in mine:
                out[(step, i)] = O.channel_block(O.KO_REAL, X, resp[i], chans[i]["shift"])[-48:].copy()

        MulticastSharder(rank, world, forward, push, lambda slot: log.append(("ready", slot)), arrive, channels_mc).run(range(nsteps))
        if rank == 0:  # every push is acknowledged before its slot's channels and before the slot is refilled
            assert [e for e in log if e[0] == "push"] == [("push", k % 2) for k in range(nsteps)]
            assert [e for e in log if e[0] == "ready"] == [("ready", k % 2) for k in range(nsteps)]
    elif os.environ.get("KA_SHARDER") == "allgather":
        # block-parallel forward: a step is `world` blocks, rank r transforms block step*world + r, ONE all-gather per
        # step hands every block's spectrum to everybody (bench.py --mg-mode allgather)
        nst = nsteps // world
        spec2 = [torch.zeros((world, N // 2 + 1), dtype=torch.complex64) for _ in range(2)]

        def forward_part(step, slot):
            spec2[slot][rank].copy_(torch.from_numpy(O.forward(O.block_window(x, L, M, step * world + rank))))

        def gather(slot):
            parts = [spec2[slot][r] for r in range(world)]
            return dist.all_gather(parts, spec2[slot][rank].clone(), async_op=True)

        def channels_ag(step, slot):
            for r in range(world):
                X = spec2[slot][r].numpy()
                for i in mine:
                    out[(step * world + r, i)] = O.channel_block(O.KO_REAL, X, resp[i], chans[i]["shift"])[-48:].copy()

        PipelinedSharder(rank, world, forward_part, gather, channels_ag, forward_on_all=True).run(range(nst))
    elif os.environ.get("KA_SHARDER", "").startswith("a2a"):
        # block-parallel forward + slice hand-off (bench.py --mg-mode a2a): rank r transforms block step*world + r and ONE
        # all-to-all gives every rank the bins ITS channels read of every block of the step.  "a2a" puts the channel groups at
        # a regular spacing (one strided copy per pack / unpack), "a2a-irregular" does not (one copy per peer).
        from ka9q_radio_b200.sharding import SliceExchange

        if os.environ["KA_SHARDER"] == "a2a":
            chans = [dict(olen=48, shift=600 + 150 * i, low=-0.3, high=0.3, beta=9.0) for i in range(8)]
        else:   # second group's window + the common width would run past the spectrum: per-peer copies
            chans = [dict(olen=48, shift=sh, low=-0.3, high=0.3, beta=9.0) for sh in (300, 700, 1100, 1500, 2700, 2800, 2900)]
        resp = {i: O.design_response(60, 48, N, True, -0.3, 0.3, 9.0) for i in range(len(chans))}
        groups = channel_groups(len(chans), world)
        mine = groups[rank]
        nb = N // 2 + 1
        stride = (nb + 3) // 4 * 4
        win = []
        for g in groups:
            sh = [chans[i]["shift"] for i in g]
            win.append((max(0, min(sh) - 32), min(nb, max(sh) + 32)))
        sx = SliceExchange(rank, world, win, stride, 1)
        assert sx.regular == (os.environ["KA_SHARDER"] == "a2a")
        nst = nsteps // world
        spec2 = [torch.full((world, stride), float("nan"), dtype=torch.complex64) for _ in range(2)]
        send = [torch.zeros((world, 1, sx.width), dtype=torch.complex64) for _ in range(2)]
        recv = [torch.zeros((world, 1, sx.width), dtype=torch.complex64) for _ in range(2)]

        def forward_part(step, slot):
            spec2[slot].fill_(float("nan"))   # a bin nobody sent would poison the channel outputs
            own = spec2[slot][rank:rank + 1]
            own[0, :nb].copy_(torch.from_numpy(O.forward(O.block_window(x, L, M, step * world + rank))))
            sx.pack(own, send[slot])

        class _Then:
            def __init__(self, h, fn):
                self.h, self.fn = h, fn

            def wait(self):
                self.h.wait()
                self.fn()

        def exchange(slot):
            h = dist.all_to_all_single(recv[slot].view(-1), send[slot].view(-1), async_op=True)
            return _Then(h, lambda: sx.unpack(spec2[slot], recv[slot]))

        def channels_a2a(step, slot):
            for r in range(world):
                X = spec2[slot][r, :nb].numpy()
                for i in mine:
                    out[(step * world + r, i)] = O.channel_block(O.KO_REAL, X, resp[i], chans[i]["shift"])[-48:].copy()

        PipelinedSharder(rank, world, forward_part, exchange, channels_a2a, forward_on_all=True).run(range(nst))
    else:
        PipelinedSharder(rank, world, forward, broadcast, channels).run(range(nsteps))
    np.save(Path(tmp) / f"rank{rank}.npy", {k: v for k, v in out.items()}, allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("sharder", ["pipelined", "multicast", "allgather", "a2a", "a2a-irregular"])
def test_two_rank_gloo_pipeline_matches_single_process(tmp_path, oracle, sharder, monkeypatch):
    import torch.multiprocessing as mp

    monkeypatch.setenv("KA_SHARDER", sharder)
    port = 29600 + (os.getpid() % 300) + {"pipelined": 0, "multicast": 17, "allgather": 31, "a2a": 43, "a2a-irregular": 59}[sharder]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = {}
    for r in range(2):
        got.update(np.load(tmp_path / f"rank{r}.npy", allow_pickle=True).item())
    L, M, nsteps = 4800, 1201, 5
    x = oracle.siggen_real(nsteps * L, 0.1, 0.02, 0.27, 1.0)
    chans = [dict(olen=48, shift=900 + 200 * i, low=-0.3, high=0.3, beta=9.0) for i in range(7)]
    if sharder == "a2a":
        chans = [dict(olen=48, shift=600 + 150 * i, low=-0.3, high=0.3, beta=9.0) for i in range(8)]
    if sharder == "a2a-irregular":
        chans = [dict(olen=48, shift=sh, low=-0.3, high=0.3, beta=9.0) for sh in (300, 700, 1100, 1500, 2700, 2800, 2900)]
    ref, _ = oracle.run_stream(x, L, M, chans)
    nexp = nsteps if sharder in ("pipelined", "multicast") else (nsteps // 2) * 2
    assert len(got) == nexp * len(chans)
    for (step, i), y in got.items():
        assert np.array_equal(y, ref[step][i])

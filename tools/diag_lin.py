"""diagnostic: F(a+b) = F(a) + F(b) at the cfg-2 size, float input, after other masters were used in the same process"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ka9q_radio_b200 import capi, workloads
from ka9q_radio_b200.channelizer import Channelizer
dev = torch.device("cuda:0")
def lin(tag):
    L, M = 2592000, 648001
    cz = Channelizer(L, M, capi.KGPU_REAL, dev, capacity=4)
    rng = np.random.default_rng(0)
    a = rng.standard_normal(L).astype(np.float32); b = rng.standard_normal(L).astype(np.float32)
    sa, sb, sab = cz.alloc_spectra(1), cz.alloc_spectra(1), cz.alloc_spectra(1)
    da, db, dab = cz.stage_stream(a), cz.stage_stream(b), cz.stage_stream(a + b)
    cz.forward(da, 1, sa); cz.forward(db, 1, sb); cz.forward(dab, 1, sab)
    torch.cuda.synchronize()
    d = (sab - sa - sb).abs()[0, :cz.master.bins]
    mx = sab.abs().max().item()
    bad = (d > 1e-5 * mx).nonzero().flatten().cpu().numpy()
    print(tag, "err", d.max().item() / mx, "bad bins", len(bad), bad[:12], bad[-5:] if len(bad) else "")
    # repeat the same input twice: determinism
    s2 = cz.alloc_spectra(1); cz.forward(dab, 1, s2); torch.cuda.synchronize()
    print(tag, "repeat identical:", torch.equal(s2, sab))
    cz.close()
lin("fresh")
w = workloads.cfg4()
cz = Channelizer(w.L, w.M, w.in_type, dev, capacity=8)
for c in w.channels[:8]: cz.add_channel(c.olen, c.shift, c.low, c.high, c.beta)
xi = w.stream(4); spec, out = cz.alloc_spectra(4), cz.alloc_outputs(4)
cz.forward(cz.stage_stream(xi), 4, spec, scale=w.scale); cz.channels(spec, 4, out); torch.cuda.synchronize(); cz.close()
lin("after cfg4")
w = workloads.cfg2(with_inverted=True)
cz = Channelizer(w.L, w.M, w.in_type, dev, capacity=len(w.channels))
for c in w.channels: cz.add_channel(c.olen, c.shift, c.low, c.high, c.beta)
xi = w.stream(40); spec, out = cz.alloc_spectra(32), cz.alloc_outputs(32)
cz.forward(cz.stage_stream(xi), 32, spec, scale=w.scale, first_block=5); cz.channels(spec, 32, out); torch.cuda.synchronize(); cz.close()
lin("after cfg2 x32")

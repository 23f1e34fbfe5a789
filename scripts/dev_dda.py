"""Dev: DDA kernel timing at config 2 on generated maps (f64 source, bit output) + parity vs the C oracle on a sample."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from oracle import c_oracle
from ppnet_b200 import ops
from ppnet_b200.synthetic import synthetic_segments
M, spm, R = 10000, 1024, 224
paths = ops.path_synthesize(0, 100, clearance=1.0, resolution=R, seed=3, pomax=24)
gen = ops.generate_maps(paths.to_bank(), 0, M, 10, 50, R, 50.0, 5.0, 1.0, seed=3, raster_inflate=2.24)
segs = synthetic_segments(M, spm, seed=8)
S = torch.from_numpy(segs).cuda()
out = {}
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("dda rc64 bits %.4f ms" % t(lambda: ops.dda_gridcheck_rc64(gen.bits, R, S, want=("bits",), out=out)))
k = 300
xy = np.ascontiguousarray(segs[:k * spm][:, [1, 0, 3, 2]].astype(np.float32))
want = c_oracle.dda_gridcheck(gen.bits[:k].cpu().numpy().view(np.uint32), R, xy, np.repeat(np.arange(k, dtype=np.int32), spm), threads=8)[0]
got = ops.unpack_bits(out["bits"], M * spm)[:k * spm].cpu().numpy()
print("mismatches", int((got != want).sum()), "positives %.3f" % got.mean())

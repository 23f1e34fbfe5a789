"""Runs one hot kernel a few times at config-2 size (for ncu): python scripts/prof_kernel.py verdict|dda|old64|old32"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from ppnet_b200 import ops
from ppnet_b200.synthetic import synthetic_segments
which = sys.argv[1]
M, spm, R = 10000, 1024, 224
rng = np.random.default_rng(3)
obs = np.zeros([M, 74, 3]); obs[..., 0] = rng.uniform(0, R, obs.shape[:2]); obs[..., 1] = rng.uniform(0, R, obs.shape[:2])
obs[..., 2] = rng.uniform(0, R / 10, obs.shape[:2])
cnt = rng.integers(40, 54, M).astype(np.int32)
segs = synthetic_segments(M, spm, seed=4)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
S, Ob, C = d(segs), d(obs), d(cnt)
S32 = d(segs[:, [1, 0, 3, 2]].astype(np.float32))
bits = ops.raster_circles_bits(Ob, C, R, 2.24)
out = {}
for _ in range(4):
    if which == "verdict":
        ops.verdict_fused(S, Ob, C, 4.48, want=("bits64", "bits32"), out=out)
    elif which == "old64":
        ops.segcheck_edage_f64(S, Ob, C, 4.48)
    elif which == "old32":
        ops.segcheck_mpnet_f32(S32, Ob, C, 4.48)
    elif which == "dda":
        ops.dda_gridcheck(bits, R, S32, want_first=False)
torch.cuda.synchronize()

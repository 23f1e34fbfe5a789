"""Device-time micro-benchmarks of the configurations that are parity cases rather than bench.py lines
(BASELINE.json configs 3 and 5) and of target-path synthesis.  CUDA events, 3 warm-ups, 10 timed launches."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from ppnet_b200 import ops

def timed(fn, n=10, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

out = {}
rng = np.random.default_rng(0)
C = 1 / 50 * 224
# ---- config 3: 4000 MPNet problems, ragged paths of 4..64 f32 waypoints, <= 50 circles
P = 4000
obs = np.zeros([P, 50, 3]); obs[..., 0] = rng.uniform(0, 224, (P, 50)); obs[..., 1] = rng.uniform(0, 224, (P, 50)); obs[..., 2] = rng.uniform(0, 9, (P, 50))
cnt = rng.integers(0, 51, P).astype(np.int32)
lens = rng.integers(4, 65, P)
wps = []
for L in lens:
    a, b = rng.uniform(5, 219, 2), rng.uniform(5, 219, 2)
    t = np.linspace(0, 1, L)[:, None]
    wps.append((a + t * (b - a) + rng.normal(0, 10, (L, 2))).astype(np.float32))
wp = torch.from_numpy(np.concatenate(wps)).cuda()
off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).cuda()
pm = torch.arange(P, dtype=torch.int32).cuda()
dobs, dcnt = torch.from_numpy(obs).cuda(), torch.from_numpy(cnt).cuda()
n_edges = int(lens.sum() - P)
ms = timed(lambda: ops.path_feasible_f32(wp, off, pm, dobs, dcnt, C))
out["config3_feasibility_check"] = {"ms": ms, "paths_per_s": P / ms * 1e3, "edges_per_s": n_edges / ms * 1e3, "paths": P, "edges": n_edges}
ms = timed(lambda: ops.lvc_f32(wp, off, pm, dobs, dcnt, C))
out["config3_lvc"] = {"ms": ms, "paths_per_s": P / ms * 1e3}
# the same edges as a flat segment batch (CSR by problem)
segs = torch.from_numpy(np.concatenate([np.concatenate([w[:-1], w[1:]], axis=1) for w in wps])).cuda()
soff = torch.from_numpy(np.concatenate([[0], np.cumsum(lens - 1)]).astype(np.int64)).cuda()
ms = timed(lambda: ops.segcheck_mpnet_f32(segs, dobs, dcnt, C, seg_off=soff))
out["config3_steerTo_flat"] = {"ms": ms, "segments_per_s": n_edges / ms * 1e3}
# ---- path synthesis (A1-A9)
for n in (100, 1000, 10000):
    ms = timed(lambda: ops.path_synthesize(0, n, clearance=1.0, seed=1), n=3, w=1)
    out["path_synthesize_%d" % n] = {"ms": ms, "paths_per_s": n / ms * 1e3}
# ---- config 5: dense 1024^2
R, S, O, M, SPM = 1024, 40, 400, 256, 4096
MS, CL, OS = 200.0, 4.0, 20.0
c_px = CL / MS * R
bank = ops.path_synthesize(0, 64, seg_num=S, clearance=CL, map_size=MS, resolution=R, seed=9, hmax=96, pomax=64).to_bank()
gen = ops.generate_maps(bank, 0, M, 4, O, R, MS, OS, CL, seed=9, raster_inflate=c_px / 2, max_tries=1 << 16)
ms = timed(lambda: ops.generate_maps(bank, 0, M, 4, O, R, MS, OS, CL, seed=9, raster_inflate=c_px / 2, max_tries=1 << 16, out=gen), n=5)
out["config5_generate_maps"] = {"ms": ms, "maps_per_s": M / ms * 1e3, "valid": int(gen.valid.sum().item()), "maps": M}
s = rng.uniform(0, R, (M * SPM, 2)); ang = rng.uniform(0, 2 * np.pi, M * SPM); ln = rng.uniform(64, 1024, M * SPM)
e = s + np.stack([np.cos(ang), np.sin(ang)], axis=1) * ln[:, None]
seg64 = torch.from_numpy(np.concatenate([s, e], axis=1)).cuda(); seg32 = seg64.float()
v = torch.empty(M * SPM, dtype=torch.uint8, device="cuda")
for name, fn in (("segcheck_f64", lambda: ops.segcheck_edage_f64(seg64, gen.obs, gen.obs_cnt, c_px, bound=float(R), out=v)),
                 ("segcheck_f32", lambda: ops.segcheck_mpnet_f32(seg32, gen.obs, gen.obs_cnt, c_px, bound=float(R), out=v)),
                 ("dda_gridcheck", lambda: ops.dda_gridcheck(gen.bits, R, seg32, want_first=False, out=v))):
    ms = timed(fn, n=5)
    out["config5_" + name] = {"ms": ms, "segments_per_s": M * SPM / ms * 1e3, "positives": float(v.float().mean().item())}
out["config5_avg_circles"] = float(gen.obs_cnt.float().mean().item())
print(json.dumps(out, indent=1))

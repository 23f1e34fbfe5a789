"""Randomised soak: CUDA verdict kernels vs the C oracle over many random configurations (resolution, circle counts,
clearance, bound, ragged CSR, segment length scales, outliers).  Not part of the test suite; run on a GPU box:
    python scripts/soak.py [rounds] [seed]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from oracle import c_oracle
from ppnet_b200 import ops

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
tot = 0
for it in range(rounds):
    R = int(rng.choice([33, 64, 224, 224, 224, 500, 1024]))
    M = int(rng.integers(1, 400))
    big = rng.random() < 0.3            # >= 8 x 148 maps: one CTA per map, so CTAs hold >= 256 segments and the verdict
    if big:                             # kernel's inner-disk grid is on
        M = int(rng.integers(1200, 2500))
    omax = int(rng.choice([1, 3, 50, 50, 74, 128, 129, 300]))
    clearance = float(rng.choice([0.0, 4.48, 13.44, 1e-9, 40.0])) * R / 224
    bound = float(rng.choice([R, R, R, 0.0, 1e9]))
    obs = np.zeros([M, omax, 3])
    obs[..., 0] = rng.uniform(-0.1 * R, 1.1 * R, (M, omax)); obs[..., 1] = rng.uniform(-0.1 * R, 1.1 * R, (M, omax))
    obs[..., 2] = rng.uniform(0, rng.choice([0.02, 0.1, 0.3]) * R, (M, omax))
    cnt = rng.integers(0, omax + 1, M).astype(np.int32)
    per = rng.integers(0, int(rng.choice([3, 40, 600, 3000])), M) if not big else rng.integers(256, 700, M)
    off = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    n = int(off[-1])
    if n == 0:
        continue
    s = rng.uniform(-0.05 * R, 1.05 * R, (n, 2))
    e = s + rng.normal(0, rng.choice([0.01, 0.07, 0.3, 1.0]) * R, (n, 2))
    segs = np.concatenate([s, e], axis=1)
    k = rng.integers(0, n, max(1, n // 200))
    segs[k, 2:] = segs[k, :2]                                   # degenerate
    k = rng.integers(0, n, max(1, n // 300))
    segs[k, rng.integers(0, 4, len(k))] = rng.choice([np.nan, np.inf, -np.inf, 1e20, -1e20], len(k))
    seg_map = np.repeat(np.arange(M, dtype=np.int32), per)
    for mode in (0, 1):
        v = ops.segcheck_edage_f64(d(segs), d(obs), d(cnt), clearance, seg_off=d(off), bound=bound, dot_mode=mode).cpu().numpy()
        w = c_oracle.segcheck_f64(segs, seg_map, obs, cnt, clearance, bound=bound, dot_mode=mode, threads=8)
        assert np.array_equal(v, w), ("f64", it, mode, R, M, omax, clearance, bound, np.nonzero(v != w)[0][:5])
    with np.errstate(over="ignore", invalid="ignore"):
        s32 = segs.astype(np.float32)
    v, st = ops.segcheck_mpnet_f32(d(s32), d(obs), d(cnt), clearance, seg_off=d(off), bound=bound, want_steer=True)
    w, wst = c_oracle.segcheck_f32(s32, seg_map, obs, cnt, clearance, bound=bound, threads=8)
    assert np.array_equal(v.cpu().numpy(), w) and np.array_equal(st.cpu().numpy(), wst), ("f32", it, R, M, omax, clearance, bound)
    if R <= 1024:
        infl = float(rng.choice([0.0, clearance / 2]))
        bits = ops.raster_circles_bits(d(obs), d(cnt), R, infl)
        wb = c_oracle.raster_circles_bits(obs, cnt, R, infl, threads=8)
        assert np.array_equal(bits.cpu().numpy().view(np.uint32), wb), ("raster", it, R)
        vd, fh = ops.dda_gridcheck(bits, R, d(s32), seg_off=d(off))
        wd, wfh = c_oracle.dda_gridcheck(wb, R, s32, seg_map, threads=8)
        assert np.array_equal(vd.cpu().numpy(), wd) and np.array_equal(fh.cpu().numpy(), wfh), ("dda", it, R)
    # round 2: the fused kernel (both flavours on one read, bytes and bit-packed), the DDA on the f64 array, the compaction
    with np.errstate(over="ignore", invalid="ignore"):
        xy = np.ascontiguousarray(segs[:, [1, 0, 3, 2]].astype(np.float32))
    cmp_mode = int(rng.integers(0, 2))
    fo = ops.verdict_fused(d(segs), d(obs), d(cnt), clearance, seg_off=d(off), bound=bound, dot_mode=int(it & 1), cmp_mode=cmp_mode,
                           want=("u8_64", "u8_32", "bits64", "bits32") if it % 3 else ("bits64", "bits32"))
    w64 = c_oracle.segcheck_f64(segs, seg_map, obs, cnt, clearance, bound=bound, dot_mode=int(it & 1), threads=8)
    w32 = c_oracle.segcheck_f32_cmp(xy, seg_map, obs, cnt, clearance, cmp_mode, bound=bound, threads=8)
    for key, want in (("u8_64", w64), ("u8_32", w32), ("bits64", w64), ("bits32", w32)):
        if key in fo:
            got = (ops.unpack_bits(fo[key], n) if key.startswith("bits") else fo[key]).cpu().numpy()
            assert np.array_equal(got, want), ("fused", key, it, R, M, omax, clearance, bound, cmp_mode, np.nonzero(got != want)[0][:5])
    if R <= 1024:
        do = ops.dda_gridcheck_rc64(bits, R, d(segs), seg_off=d(off), want=("bits", "u8", "first"))
        wd2, wf2 = c_oracle.dda_gridcheck(wb, R, xy, seg_map, threads=8)
        assert np.array_equal(do["u8"].cpu().numpy(), wd2) and np.array_equal(do["first"].cpu().numpy(), wf2), ("dda64", it, R)
        assert np.array_equal(ops.unpack_bits(do["bits"], n).cpu().numpy(), wd2), ("dda64 bits", it, R)
        idx, k, _ = ops.compact_bits(fo["bits64"], fo["bits32"], do["bits"], n=n)
        free = np.nonzero((w64 | w32 | wd2) == 0)[0]
        assert int(k.item()) == len(free) and np.array_equal(idx[:len(free)].cpu().numpy(), free), ("compact", it)
    tot += n
    print("round %d ok: R=%d M=%d omax=%d n=%d c=%.3g bound=%.3g positives %.2f" % (it, R, M, omax, n, clearance, bound, w.mean()), flush=True)
print("soak ok:", tot, "segments x {f64 fused / un-fused, f32 + steer, raster + DDA, fused A11+A12 (both cmp modes), DDA on the f64 array, compaction} bit-exact")

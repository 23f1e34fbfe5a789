"""Synthetic inputs of the reference's shapes for benches and size-independent tests (no datasets are
reachable).  Pure numpy/scipy host code: it prepares INPUTS, it is not part of the measured path."""
import numpy as np


def synthetic_bank(n_bank=100, n_points=1000, n_seg=10, resolution=224, pomax=24, seed=0):
    """A bank of smooth target paths in the normalised layout Path.space_normalization leaves behind
    (EDaGe-PP/Path.py:157-189): PathPoint = integer grid cells + one shared fractional offset, hull centred on
    (R/2, R/2), SegPointImage every n_points/n_seg points, a few path-hugging obstacles [x, y, r].
    Returns dict of numpy arrays: pathpt [B,Np,2], segpt [B,S+1,2], hull [B,hmax,2], hull_cnt, obs [B,pomax,3], obs_cnt."""
    from scipy.spatial import ConvexHull

    rng = np.random.default_rng(seed)
    R = float(resolution)
    hmax = 48
    pathpt = np.zeros([n_bank, n_points, 2])
    segpt = np.zeros([n_bank, n_seg + 1, 2])
    hull = np.zeros([n_bank, hmax, 2])
    hull_cnt = np.zeros(n_bank, dtype=np.int32)
    obs = np.zeros([n_bank, pomax, 3])
    obs_cnt = np.zeros(n_bank, dtype=np.int32)
    for b in range(n_bank):
        # heading = smooth random process; arc length ~ U(0.35, 0.8) R in total
        k = rng.normal(0, 1, n_seg)
        curv = np.repeat(k, n_points // n_seg) * rng.uniform(0.002, 0.012)
        heading = np.cumsum(curv) + rng.uniform(0, 2 * np.pi)
        step = rng.uniform(0.35, 0.8) * R / n_points
        xy = np.cumsum(np.stack([np.cos(heading), np.sin(heading)], axis=1) * step, axis=0)
        cells = np.rint(xy / 1.0).astype(np.int64)
        hv = ConvexHull(cells.astype(np.float64)).vertices
        centre = cells[hv].astype(np.float64).mean(axis=0)
        shift = np.array([R / 2, R / 2]) - centre
        pts = cells + shift                                   # integer grid + shared fractional offset
        pathpt[b] = pts
        idx = np.minimum(np.arange(n_seg + 1) * (n_points // n_seg), n_points - 1)
        segpt[b] = pts[idx]
        h = cells[hv] + shift
        hull[b, :len(h)] = h
        hull_cnt[b] = len(h)
        # path-hugging obstacles: circles beside the path with clearance >= 4.48 px + r to every odd point
        n_o = int(rng.integers(4, pomax // 2))
        kept = 0
        for _ in range(4 * n_o):
            if kept >= n_o:
                break
            i = int(rng.integers(50, n_points - 50))
            t = pts[min(i + 5, n_points - 1)] - pts[max(i - 5, 0)]
            nrm = np.array([t[1], -t[0]]) / (np.linalg.norm(t) + 1e-9)
            r = rng.uniform(2, 9)
            cen = pts[i] + nrm * (r + rng.uniform(6, 14)) * rng.choice([-1, 1])
            d = np.sqrt(((pts[1::2] - cen) ** 2).sum(axis=1)).min()
            if d > r + 4.48:
                obs[b, kept] = [cen[1], cen[0], r]            # [x, y, r] = [col, row, r]
                kept += 1
        obs_cnt[b] = kept
    return dict(pathpt=pathpt, segpt=segpt, hull=hull, hull_cnt=hull_cnt, obs=obs, obs_cnt=obs_cnt)


def synthetic_segments(n_maps, segs_per_map, resolution=224, sigma=15.0, seed=0):
    """SURVEY 8(d) config 2 segments: s ~ U(0,R)^2, e = s + N(0, sigma^2).  Returns f64 [N,4] (s0,s1,e0,e1)."""
    rng = np.random.default_rng(seed)
    s = rng.uniform(0, resolution, (n_maps * segs_per_map, 2))
    e = s + rng.normal(0, sigma, s.shape)
    return np.concatenate([s, e], axis=1)

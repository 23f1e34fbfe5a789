"""How many (segment, circle) pairs survive the separable bin culling on the config-2 geometry, against the number of
pairs that are truly near (distance to the segment < thr)?  numpy only.  32 bins (what the kernel uses): 1.86 candidates
per segment; 64 bins: 1.61; 128 bins: 1.44; truly near: 1.01.  DESIGN.md 4.2 uses these figures for the instruction
lower bound of the verdict kernel."""
import numpy as np
rng = np.random.default_rng(0)
R, M, spm, O = 224.0, 200, 256, 47
ox, oy, r = rng.uniform(0, R, (M, O)), rng.uniform(0, R, (M, O)), rng.uniform(0, 22.4, (M, O))
thr = r + 2.24
s = rng.uniform(0, R, (M, spm, 2)); e = s + rng.normal(0, 15, (M, spm, 2))

def cand(bins, piece_bins):
    sc = bins / R
    cx0, cx1 = np.floor(np.clip((ox - thr) * sc, 0, bins - 1)), np.floor(np.clip((ox + thr) * sc, 0, bins - 1))
    cy0, cy1 = np.floor(np.clip((oy - thr) * sc, 0, bins - 1)), np.floor(np.clip((oy + thr) * sc, 0, bins - 1))
    sx, sy, dx, dy = s[..., 0] * sc, s[..., 1] * sc, (e[..., 0] - s[..., 0]) * sc, (e[..., 1] - s[..., 1]) * sc
    npc = np.minimum(16, 1 + (np.maximum(abs(dx), abs(dy)) / piece_bins).astype(int))
    hit = np.zeros((M, spm, O), bool)
    for pc in range(16):
        act = npc > pc
        if not act.any(): break
        ta, tb = pc / npc, (pc + 1) / npc
        ax, bx, ay, by = sx + dx * ta, sx + dx * tb, sy + dy * ta, sy + dy * tb
        x0, x1 = np.floor(np.clip(np.minimum(ax, bx), 0, bins - 1)), np.floor(np.clip(np.maximum(ax, bx), 0, bins - 1))
        y0, y1 = np.floor(np.clip(np.minimum(ay, by), 0, bins - 1)), np.floor(np.clip(np.maximum(ay, by), 0, bins - 1))
        hit |= ((cx1[:, None, :] >= x0[..., None]) & (cx0[:, None, :] <= x1[..., None]) & (cy1[:, None, :] >= y0[..., None]) &
                (cy0[:, None, :] <= y1[..., None]) & act[..., None])
    return hit.sum() / (M * spm)

d = e - s; L2 = (d ** 2).sum(-1)
q = np.stack([ox[:, None, :] - s[..., 0:1], oy[:, None, :] - s[..., 1:2]], -1)
t = np.clip((q * d[:, :, None, :]).sum(-1) / L2[..., None], 0, 1)
pr = s[:, :, None, :] + t[..., None] * d[:, :, None, :]
dist = np.sqrt((ox[:, None, :] - pr[..., 0]) ** 2 + (oy[:, None, :] - pr[..., 1]) ** 2)
print("truly near pairs per segment %.2f" % ((dist < thr[:, None, :]).sum() / (M * spm)))
for bins, pb in ((32, 12), (64, 24), (128, 48)):
    print("%3d bins: %.2f candidates per segment" % (bins, cand(bins, pb)))

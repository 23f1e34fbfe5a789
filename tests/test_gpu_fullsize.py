"""GPU tests at BASELINE.json's full sizes through size-independent properties (the oracle only sees a sample):
config 2 (10 k maps x 1024 segments), config 4 (sharded generation, counts only cross ranks), config 5 (dense
1024^2 maps, 400 circles, 40-piece paths, 4096 long segments per map)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import philox

pytestmark = pytest.mark.gpu

C = 1 / 50 * 224


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "these tests need a B200"
    from ppnet_b200 import ops as _ops
    return _ops


def _checksum(t):
    """Order-sensitive 64-bit checksum of a tensor's bytes, computed on the device."""
    b = t.contiguous().view(torch.uint8).reshape(-1).to(torch.int64)
    w = (torch.arange(b.numel(), device=b.device, dtype=torch.int64) % 1000003) + 1
    return int((b * w).sum().item())


def test_config2_full_size_properties(ops):
    from ppnet_b200.synthetic import synthetic_segments
    M, SPM, R, O = 10000, 1024, 224, 50
    paths = ops.path_synthesize(0, 100, clearance=1.0, resolution=R, seed=3, pomax=24)
    bank = paths.to_bank()
    gen = ops.generate_maps(bank, 0, M, 10, O, R, 50.0, 5.0, 1.0, seed=3, raster_inflate=C / 2)
    assert int(gen.valid.sum().item()) == M
    s64 = torch.from_numpy(synthetic_segments(M, SPM, seed=8)).cuda()
    s32 = s64.to(torch.float32)
    v64 = ops.segcheck_edage_f64(s64, gen.obs, gen.obs_cnt, C)
    v32 = ops.segcheck_mpnet_f32(s32, gen.obs, gen.obs_cnt, C)
    vd = ops.dda_gridcheck(gen.bits, R, s32, want_first=False)
    # (1) the verdict is an OR over the circles: any order of a map's circles gives the same bytes
    omax = gen.obs.shape[1]
    rank_ok = torch.arange(omax, device="cuda")[None, :] < gen.obs_cnt[:, None]
    key = torch.where(rank_ok, torch.rand([M, omax], device="cuda"), torch.full([M, omax], 2.0, device="cuda"))
    perm = torch.argsort(key, dim=1)                                     # valid rows shuffled, padding stays behind
    obs_p = torch.gather(gen.obs, 1, perm[:, :, None].expand(-1, -1, 3)).contiguous()
    assert torch.equal(ops.segcheck_edage_f64(s64, obs_p, gen.obs_cnt, C), v64)
    assert torch.equal(ops.segcheck_mpnet_f32(s32, obs_p, gen.obs_cnt, C), v32)
    # (2) sharding the maps (ragged split) changes nothing
    cut = [0, 3333, 7001, M]
    parts = [ops.segcheck_edage_f64(s64[a * SPM:b * SPM], gen.obs[a:b].contiguous(), gen.obs_cnt[a:b].contiguous(), C)
             for a, b in zip(cut, cut[1:])]
    assert torch.equal(torch.cat(parts), v64)
    # (3) the verdict byte and the first blocked step tell the same story; a zero-length segment is blocked iff its cell is
    vd2, fh = ops.dda_gridcheck(gen.bits, R, s32)
    assert torch.equal(vd2, vd) and torch.equal(vd != 0, fh >= 0)
    pts = torch.cat([s32[:, :2], s32[:, :2]], dim=1).contiguous()
    vp, fp = ops.dda_gridcheck(gen.bits, R, pts)
    assert torch.equal((fp == 0), vp != 0)                              # blocked at k = 0 or free, nothing else
    assert torch.equal((fh == 0), vp != 0)                              # and that is the first cell of the real walk
    # (5) the oracle on EVERY map: all 3 x 10.24 M verdicts (the C oracle does ~5e7 verdicts/s on the box's host threads)
    import os
    th = os.cpu_count() or 4
    obs_h, cnt_h = gen.obs.cpu().numpy(), gen.obs_cnt.cpu().numpy()
    seg_h = s64.cpu().numpy()
    sm = np.repeat(np.arange(M, dtype=np.int32), SPM)
    assert np.array_equal(v64.cpu().numpy(), c_oracle.segcheck_f64(seg_h, sm, obs_h, cnt_h, C, threads=th))
    seg32_h = s32.cpu().numpy()
    assert np.array_equal(v32.cpu().numpy(), c_oracle.segcheck_f32(seg32_h, sm, obs_h, cnt_h, C, threads=th, want_steer=False))
    bits_h = gen.bits.cpu().numpy().view(np.uint32)
    assert np.array_equal(vd.cpu().numpy(), c_oracle.dda_gridcheck(bits_h, R, seg32_h, sm, threads=th)[0])
    # (6) the round-2 hot path at full size: ONE f64 array -> fused verdicts + DDA, bit-packed, survivors compacted
    xy = np.ascontiguousarray(seg_h[:, [1, 0, 3, 2]].astype(np.float32))
    w32 = c_oracle.segcheck_f32(xy, sm, obs_h, cnt_h, C, threads=th, want_steer=False)
    wdd = c_oracle.dda_gridcheck(bits_h, R, xy, sm, threads=th)[0]
    fo = ops.verdict_fused(s64, gen.obs, gen.obs_cnt, C, want=("bits64", "bits32"))
    do = ops.dda_gridcheck_rc64(gen.bits, R, s64, want=("bits",))
    n = M * SPM
    assert torch.equal(ops.unpack_bits(fo["bits64"], n), v64)
    assert np.array_equal(ops.unpack_bits(fo["bits32"], n).cpu().numpy(), w32)
    assert np.array_equal(ops.unpack_bits(do["bits"], n).cpu().numpy(), wdd)
    idx, k, _ = ops.compact_bits(fo["bits64"], fo["bits32"], do["bits"], n=n)
    free = np.nonzero((v64.cpu().numpy() | w32 | wdd) == 0)[0]
    assert int(k.item()) == len(free) and np.array_equal(idx[:len(free)].cpu().numpy(), free)
    assert 0.2 < float(v64.float().mean()) < 0.8


def test_config4_sharded_generation_counts_and_bytes(ops):
    """1 launch of 200 k maps == 4 contiguous shards (what 4 ranks would produce): identical labels / obstacle sets /
    bitmaps (checksums), and the four counters add up -- the only numbers that cross ranks."""
    from ppnet_b200 import sharding
    R, O, total = 224, 50, 200000
    bank = ops.path_synthesize(0, 1000, clearance=1.0, resolution=R, seed=5, pomax=24).to_bank()
    cnt_all = torch.zeros(4, dtype=torch.int64, device="cuda")
    whole = ops.generate_maps(bank, 0, total, 10, O, R, 50.0, 5.0, 1.0, seed=5, counters=cnt_all, want_labels=False)
    sums = {k: _checksum(getattr(whole, k)) for k in ("angle", "trans", "obs_cnt", "rand_cnt", "tries", "bits")}
    obs_sum = float(whole.obs.sum().item())
    n_valid = int(whole.valid.sum().item())
    del whole
    torch.cuda.empty_cache()
    per_rank = []
    acc = {k: [] for k in sums}
    obs_parts = 0.0
    for r in range(4):
        first, count = sharding.shard_range(total, r, 4)
        c = torch.zeros(4, dtype=torch.int64, device="cuda")
        part = ops.generate_maps(bank, first, count, 10, O, R, 50.0, 5.0, 1.0, seed=5, counters=c, want_labels=False)
        per_rank.append(c.cpu())
        for k in acc:
            acc[k].append(getattr(part, k))
        obs_parts += float(part.obs.sum().item())
    for k in sums:
        assert _checksum(torch.cat(acc[k])) == sums[k], k
    assert abs(obs_parts - obs_sum) <= 1e-9 * abs(obs_sum)
    tot = torch.stack(per_rank).sum(0)
    assert torch.equal(tot, cnt_all.cpu())
    assert int(tot[0]) == total and int(tot[1]) == n_valid == total


def test_config5_dense_1024_long_paths(ops):
    """R = 1024 (c_px = 20.48), 40-piece target paths (4000 points), 400 candidate circles per map, 4096 long
    segments per map: multi-tile circle staging, 131 KB bitmaps through the bulk-copy path, 160 KB generator CTAs."""
    R, S, O, M, SPM = 1024, 40, 400, 6, 4096
    MS, CL, OS = 200.0, 4.0, 20.0                              # a 40-piece path needs a 200-unit map; same pixel scale:
    c_px = CL / MS * R                                         # c_px = 20.48, r_px ~ U(0, 102.4)
    paths = ops.path_synthesize(0, 3, seg_num=S, clearance=CL, map_size=MS, resolution=R, seed=9, hmax=96, pomax=64)
    assert int(paths.hull_cnt.max().item()) <= 96
    bank = paths.to_bank()
    gen = ops.generate_maps(bank, 0, M, 2, O, R, MS, OS, CL, seed=9, raster_inflate=c_px / 2, max_tries=1 << 16)
    torch.cuda.synchronize()
    assert int(gen.valid.sum().item()) == M
    # A14 against the oracle on the same Philox candidates and the device's own labels
    pp = gen.pathpt.cpu().numpy()
    cand = np.stack([philox.candidates(9, g, O, MS, OS) for g in range(M)])
    w_acc, w_out, w_cnt = c_oracle.clearance_filter(pp, cand, MS, float(R), CL, threads=4)
    rc = gen.rand_cnt.cpu().numpy()
    assert np.array_equal(rc, w_cnt)
    obs = gen.obs.cpu().numpy()
    for m in range(M):
        assert np.array_equal(obs[m, :rc[m]], w_out[m, :rc[m]])
    cnt = gen.obs_cnt.cpu().numpy()
    assert cnt.max() > 128                                     # more than one 128-circle tile in the verdict kernels
    # raster at 1024^2 == oracle raster of the same obstacle sets
    bits = gen.bits.cpu().numpy().view(np.uint32)
    assert np.array_equal(bits, c_oracle.raster_circles_bits(obs, cnt, R, c_px / 2, threads=4))
    # long segments, 64..1024 px
    rng = np.random.default_rng(4)
    s = rng.uniform(0, R, (M * SPM, 2))
    ang = rng.uniform(0, 2 * np.pi, M * SPM)
    ln = rng.uniform(64, 1024, M * SPM)
    e = s + np.stack([np.cos(ang), np.sin(ang)], axis=1) * ln[:, None]
    segs = np.concatenate([s, e], axis=1)
    sm = np.repeat(np.arange(M, dtype=np.int32), SPM)
    d64, d32 = torch.from_numpy(segs).cuda(), torch.from_numpy(segs.astype(np.float32)).cuda()
    v64 = ops.segcheck_edage_f64(d64, gen.obs, gen.obs_cnt, c_px, bound=float(R)).cpu().numpy()
    assert np.array_equal(v64, c_oracle.segcheck_f64(segs, sm, obs, cnt, c_px, bound=float(R), threads=8))
    v32, st = ops.segcheck_mpnet_f32(d32, gen.obs, gen.obs_cnt, c_px, bound=float(R), want_steer=True)
    w32, wst = c_oracle.segcheck_f32(segs.astype(np.float32), sm, obs, cnt, c_px, bound=float(R), threads=8)
    assert np.array_equal(v32.cpu().numpy(), w32) and np.array_equal(st.cpu().numpy(), wst)
    vd, fh = ops.dda_gridcheck(gen.bits, R, d32)
    wd, wfh = c_oracle.dda_gridcheck(bits, R, segs.astype(np.float32), sm, threads=8)
    assert np.array_equal(vd.cpu().numpy(), wd) and np.array_equal(fh.cpu().numpy(), wfh)
    assert 0.3 < v64.mean() <= 1.0 and int(fh.max().item()) > 100
    # the same resolution, NON-degenerate: 400 small circles (r <= 4 px, clearance 2 px) and segments of 64..448 px, so that
    # about half of the segments are free and walk their whole length through the 131 KB bitmap (hundreds of cells)
    CL2, OS2 = 0.4, 0.8
    c2 = CL2 / MS * R
    # (unchecked synthesis: with a 2 px clearance set_obstacles hits its iteration guard on some isles -- the reference would
    # spin there -- and the path-hugging obstacles do not matter for this workload)
    paths2 = ops.path_synthesize(0, 3, seg_num=S, clearance=CL2, map_size=MS, resolution=R, seed=9, hmax=128, pomax=64)
    assert int(paths2.hull_cnt.max().item()) <= 128
    gen2 = ops.generate_maps(paths2.to_bank(), 0, M, 2, O, R, MS, OS2, CL2, seed=9, raster_inflate=c2 / 2, max_tries=1 << 16)
    assert int(gen2.valid.sum().item()) == M
    ln2 = rng.uniform(64, 448, M * SPM)
    e2 = s + np.stack([np.cos(ang), np.sin(ang)], axis=1) * ln2[:, None]
    segs2 = np.concatenate([s, e2], axis=1)
    xy2 = np.ascontiguousarray(segs2[:, [1, 0, 3, 2]].astype(np.float32))
    obs2, cnt2 = gen2.obs.cpu().numpy(), gen2.obs_cnt.cpu().numpy()
    bits2 = gen2.bits.cpu().numpy().view(np.uint32)
    D2 = torch.from_numpy(segs2).cuda()
    fo = ops.verdict_fused(D2, gen2.obs, gen2.obs_cnt, c2, bound=float(R), want=("u8_64", "u8_32"))
    w64 = c_oracle.segcheck_f64(segs2, sm, obs2, cnt2, c2, bound=float(R), threads=8)
    w32 = c_oracle.segcheck_f32(xy2, sm, obs2, cnt2, c2, bound=float(R), threads=8, want_steer=False)
    assert np.array_equal(fo["u8_64"].cpu().numpy(), w64) and np.array_equal(fo["u8_32"].cpu().numpy(), w32)
    do = ops.dda_gridcheck_rc64(gen2.bits, R, D2, want=("u8", "first"))
    wd2, wf2 = c_oracle.dda_gridcheck(bits2, R, xy2, sm, threads=8)
    assert np.array_equal(do["u8"].cpu().numpy(), wd2) and np.array_equal(do["first"].cpu().numpy(), wf2)
    assert 0.3 < w64.mean() < 0.7 and 0.3 < wd2.mean() < 0.8
    free_walk = np.maximum(np.abs(np.rint(xy2[:, 2]) - np.rint(xy2[:, 0])), np.abs(np.rint(xy2[:, 3]) - np.rint(xy2[:, 1])))[wd2 == 0]
    assert free_walk.mean() > 150                              # the free segments really are long walks

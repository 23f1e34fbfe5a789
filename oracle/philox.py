"""TEST INFRASTRUCTURE ONLY -- numpy Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as
1, 2, 3", SC'11; Random123 reference constants) and the draw layouts of the B200 generator kernels.
Known-answer vectors (Random123 kat_vectors) are checked in tests/test_philox_cpu.py."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_PLACE, STREAM_OBST, STREAM_GMM_PARAM, STREAM_GMM_SAMPLE, STREAM_UNIFORM, STREAM_PATH = 1, 2, 3, 4, 5, 6
STREAM_PATH_OBST = 7


def philox4x32_10(key, ctr):
    """key (k0, k1) ints; ctr uint32[n,4] -> uint32[n,4]."""
    c = np.asarray(ctr, dtype=np.uint64).reshape(-1, 4).copy()
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[:, 0]
        p1 = M1 * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c[:, 1] ^ np.uint64(k0)
        n2 = hi0 ^ c[:, 3] ^ np.uint64(k1)
        c = np.stack([n0, lo1, n2, lo0], axis=1)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def key_of(seed):
    return (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def u53(a, b):
    """53-bit double in [0,1) from two words (numpy random_sample construction)."""
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    return ((a >> np.uint64(5)) * np.uint64(1 << 26) + (b >> np.uint64(6))).astype(np.float64) * (1.0 / 9007199254740992.0)


def u24(a):
    return (np.asarray(a, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def _ctr(block, stream, unit):
    block = np.asarray(block, dtype=np.uint64).reshape(-1)
    unit = np.broadcast_to(np.asarray(unit, dtype=np.uint64), block.shape)
    return np.stack([block, np.full(block.shape, stream, dtype=np.uint64), unit & MASK, unit >> np.uint64(32)], axis=1)


def uniform(seed, stream, unit0, n_units, per_unit):
    """ppnet_uniform_f64: out[u][k]; block = k//2, words (x,y) for even k, (z,w) for odd k."""
    out = np.empty([n_units, per_unit])
    nb = (per_unit + 1) // 2
    for u in range(n_units):
        r = philox4x32_10(key_of(seed), _ctr(np.arange(nb), stream, unit0 + u))
        vals = np.stack([u53(r[:, 0], r[:, 1]), u53(r[:, 2], r[:, 3])], axis=1).reshape(-1)
        out[u] = vals[:per_unit]
    return out


def placement_draw(seed, g, t, resolution):
    """Try t of map g -> (angle_deg, t0, t1):  MapGenerate.py:63-64 applied to Philox uniforms."""
    r = philox4x32_10(key_of(seed), _ctr([2 * t, 2 * t + 1], STREAM_PLACE, g))
    R = np.float64(resolution)
    angle = u53(r[0, 0], r[0, 1]) * 360 - 180
    t0 = int(u53(r[0, 2], r[0, 3]) * R - R / 2)
    t1 = int(u53(r[1, 0], r[1, 1]) * R - R / 2)
    return float(angle), t0, t1


def candidates(seed, g, O, map_size, obstacle_size):
    """cand[O,3] = (x, y, r) map units: block j -> (x, y); block O + j//2, half j&1 -> r."""
    a = philox4x32_10(key_of(seed), _ctr(np.arange(O), STREAM_OBST, g))
    b = philox4x32_10(key_of(seed), _ctr(O + np.arange((O + 1) // 2), STREAM_OBST, g))
    x = u53(a[:, 0], a[:, 1]) * np.float64(map_size)
    y = u53(a[:, 2], a[:, 3]) * np.float64(map_size)
    rr = np.stack([u53(b[:, 0], b[:, 1]), u53(b[:, 2], b[:, 3])], axis=1).reshape(-1)[:O]
    return np.stack([x, y, rr * np.float64(obstacle_size)], axis=1)


def gmm_params(seed, order, dim, mean_range, std_range):
    r = philox4x32_10(key_of(seed), _ctr(np.arange(order * dim), STREAM_GMM_PARAM, 0))
    mean = (u24(r[:, 0]) * np.float32(mean_range)).reshape(order, dim)
    std = (u24(r[:, 1]) * np.float32(std_range)).reshape(order, dim)
    w = u24(r[:order, 2])
    return mean, std, w


def gmm_sample(seed, sample0, n, mean, std, w):
    """Layout (ppnet_b200/csrc/rng.cu): block 0 of sample i: x -> component, (y, z) -> Box-Muller pair of dims 0/1;
    dims 2q+2, 2q+3 (q >= 0) take block 1 + q // 2, words (x, y) for even q and (z, w) for odd q.  float32 math; the
    device uses the SFU __logf / __sincosf / sqrt.approx, so values agree to ~1e-4 absolute, components exactly."""
    K, D = mean.shape
    idx = sample0 + np.arange(n, dtype=np.uint64)
    r = philox4x32_10(key_of(seed), _ctr(np.zeros(n), STREAM_GMM_SAMPLE, idx))
    tot = np.float32(0)
    for k in range(K):
        tot = np.float32(tot + w[k])
    cdf = np.empty(K, dtype=np.float32)
    acc = np.float32(0)
    for k in range(K):
        acc = np.float32(acc + w[k])
        cdf[k] = np.float32(acc / tot)
    uc = u24(r[:, 0])
    comp = np.minimum((uc[:, None] >= cdf[None, :K - 1]).sum(axis=1), K - 1) if K > 1 else np.zeros(n, dtype=np.int64)
    # `while k < K-1 and uc >= cdf[k]` stops at the first k with uc < cdf[k]; cdf is non-decreasing
    comp = np.asarray([next((k for k in range(K - 1) if not uc[i] >= cdf[k]), K - 1) for i in range(n)]) if n <= 4096 else comp

    def pair(wa, wb):
        u1 = ((wa >> np.uint32(8)).astype(np.float32) + np.float32(1)) * np.float32(1.0 / 16777216.0)
        u2 = u24(wb)
        rad = np.sqrt(np.float32(-2) * np.log(u1)).astype(np.float32)
        return (rad * np.cos(np.float32(2 * np.pi) * u2).astype(np.float32),
                rad * np.sin(np.float32(2 * np.pi) * u2).astype(np.float32))

    out = np.empty([n, D], dtype=np.float32)
    blk = None
    for d in range(0, D, 2):
        if d == 0:
            z0, z1 = pair(r[:, 1], r[:, 2])
        else:
            q = (d - 2) >> 1
            if q % 2 == 0:
                blk = philox4x32_10(key_of(seed), _ctr(np.full(n, 1 + q // 2), STREAM_GMM_SAMPLE, idx))
            z0, z1 = pair(blk[:, 2], blk[:, 3]) if q % 2 else pair(blk[:, 0], blk[:, 1])
        out[:, d] = mean[comp, d] + std[comp, d] * z0
        if d + 1 < D:
            out[:, d + 1] = mean[comp, d + 1] + std[comp, d + 1] * z1
    return out, comp.astype(np.int32)


def path_draws(seed, g, seg_num):
    """A1 draws of path g (ppnet_b200/csrc/path_synth.cu): block 0 (x,y) -> forced-straight draw; piece i: block
    1 + 502 i: (x,y) -> is_straight draw, (z,w) -> EndPoint draw; blocks 2 + 502 i + j (j < 500): y[2j], y[2j+1].
    -> (forced bool, straight bool[S], y f64[S,1000], u_end f64[S])."""
    key = key_of(seed)
    r0 = philox4x32_10(key, _ctr([0], STREAM_PATH, g))
    forced = not (u53(r0[0, 0], r0[0, 1]) > 0.01)
    straight, ys, ue = [], [], []
    for i in range(seg_num):
        b0 = 1 + 502 * i
        r = philox4x32_10(key, _ctr(b0 + np.arange(501), STREAM_PATH, g))
        straight.append(forced or bool(u53(r[0, 0], r[0, 1]) < 0.2))
        ue.append(float(u53(r[0, 2], r[0, 3])))
        ys.append(np.stack([u53(r[1:, 0], r[1:, 1]), u53(r[1:, 2], r[1:, 3])], axis=1).reshape(-1))
    return forced, np.asarray(straight), np.asarray(ys), np.asarray(ue)


def path_obst_draws(seed, g, n):
    """A9 draws of path g: draw t = u24 of word (t & 3) of block (t >> 2), STREAM_PATH_OBST."""
    nb = (n + 3) // 4
    r = philox4x32_10(key_of(seed), _ctr(np.arange(nb), STREAM_PATH_OBST, g))
    return u24(r.reshape(-1))[:n]


STREAM_SEGS = 8


def propose_segments(seed, g, segs_per_map, resolution, sigma):
    """Candidate segments of global map g (ppnet_b200/csrc/rng.cu propose_segments_kernel): segment k: block 2k ->
    start (s_row, s_col) = u53 * R; block 2k+1 -> Box-Muller offset in float64, u1 = u53 + 2^-53 in (0, 1].
    -> f64[spm, 4] (s_row, s_col, e_row, e_col).  The starts are bit-exact; the ends go through log / sincospi, which
    differ from libm in the last ulp (tests compare them to ~1e-12 relative)."""
    k = np.arange(segs_per_map)
    a = philox4x32_10(key_of(seed), _ctr(2 * k, STREAM_SEGS, g))
    b = philox4x32_10(key_of(seed), _ctr(2 * k + 1, STREAM_SEGS, g))
    R = np.float64(resolution)
    s0, s1 = u53(a[:, 0], a[:, 1]) * R, u53(a[:, 2], a[:, 3]) * R
    u1 = u53(b[:, 0], b[:, 1]) + 1.0 / 9007199254740992.0
    rad = np.sqrt(-2.0 * np.log(u1)) * np.float64(sigma)
    ang = 2.0 * u53(b[:, 2], b[:, 3])
    return np.stack([s0, s1, s0 + rad * np.cos(np.pi * ang), s1 + rad * np.sin(np.pi * ang)], axis=1)

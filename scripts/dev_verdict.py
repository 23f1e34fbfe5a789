"""Dev check on a B200: fused A11+A12 verdict kernel vs the one-flavour kernels vs the C oracle, with timings."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from oracle import c_oracle
from ppnet_b200 import ops
from ppnet_b200.synthetic import synthetic_segments

def make(M, spm, O=50, seed=0, R=224.0):
    rng = np.random.default_rng(seed)
    obs = np.zeros([M, O + 24, 3])
    obs[..., 0] = rng.uniform(0, R, obs.shape[:2]); obs[..., 1] = rng.uniform(0, R, obs.shape[:2])
    obs[..., 2] = rng.uniform(0, R / 10, obs.shape[:2])
    cnt = rng.integers(40, 54, M).astype(np.int32)
    segs = synthetic_segments(M, spm, seed=seed + 1)
    return segs, obs, cnt

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
clear = 4.48
# ---- correctness at a size the oracle finishes quickly
M, spm = 512, 1024
segs, obs, cnt = make(M, spm)
seg_map = np.repeat(np.arange(M, dtype=np.int32), spm)
xy32 = np.ascontiguousarray(segs[:, [1, 0, 3, 2]].astype(np.float32))
w64 = c_oracle.segcheck_f64(segs, seg_map, obs, cnt, clear, threads=16)
w32 = c_oracle.segcheck_f32(xy32, seg_map, obs, cnt, clear, threads=16, want_steer=False)
w32c = c_oracle.segcheck_f32_cmp(xy32, seg_map, obs, cnt, clear, 1, threads=16)
for want in (("bits64", "bits32"), ("u8_64", "u8_32"), ("u8_64", "u8_32", "bits64", "bits32"), ("bits64",), ("u8_32",)):
    o = ops.verdict_fused(d(segs), d(obs), d(cnt), clear, want=want)
    for k, v in o.items():
        got = (ops.unpack_bits(v, len(segs)) if k.startswith("bits") else v).cpu().numpy()
        ref = w64 if k.endswith("64") else w32
        print(want, k, "mismatches", int((got != ref).sum()), "positives %.3f" % got.mean())
o = ops.verdict_fused(d(segs), d(obs), d(cnt), clear, want=("u8_32",), cmp_mode=1)
print("cmp_mode numpy1 mismatches", int((o["u8_32"].cpu().numpy() != w32c).sum()), "differs from nep50 in", int((w32c != w32).sum()))
# ragged CSR (unaligned words)
rng = np.random.default_rng(5)
lens = rng.integers(0, 300, M); lens[3] = 0
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
n = int(off[-1]); sm2 = np.repeat(np.arange(M, dtype=np.int32), lens)
o = ops.verdict_fused(d(segs[:n]), d(obs), d(cnt), clear, seg_off=d(off), want=("u8_64", "u8_32", "bits64", "bits32"))
r64 = c_oracle.segcheck_f64(segs[:n], sm2, obs, cnt, clear, threads=16)
r32 = c_oracle.segcheck_f32(xy32[:n], sm2, obs, cnt, clear, threads=16, want_steer=False)
for k, v in o.items():
    got = (ops.unpack_bits(v, n) if k.startswith("bits") else v).cpu().numpy()
    print("csr", k, "mismatches", int((got != (r64 if k.endswith("64") else r32)).sum()))
# many circles (3 tiles)
segs5, obs5, _ = make(64, 2048, O=376, seed=9)
cnt5 = np.full(64, 390, dtype=np.int32); sm5 = np.repeat(np.arange(64, dtype=np.int32), 2048)
xy5 = np.ascontiguousarray(segs5[:, [1, 0, 3, 2]].astype(np.float32))
o = ops.verdict_fused(d(segs5), d(obs5), d(cnt5), clear, want=("u8_64", "u8_32", "bits64", "bits32"))
r64 = c_oracle.segcheck_f64(segs5, sm5, obs5, cnt5, clear, threads=16); r32 = c_oracle.segcheck_f32(xy5, sm5, obs5, cnt5, clear, threads=16, want_steer=False)
for k, v in o.items():
    got = (ops.unpack_bits(v, len(segs5)) if k.startswith("bits") else v).cpu().numpy()
    print("tiles", k, "mismatches", int((got != (r64 if k.endswith("64") else r32)).sum()), "pos %.3f" % got.mean())
o = ops.verdict_fused(d(segs5), d(obs5), d(cnt5), clear, want=("bits64", "bits32"))
for k, v in o.items():
    got = ops.unpack_bits(v, len(segs5)).cpu().numpy()
    print("tiles bits-only", k, "mismatches", int((got != (r64 if k.endswith("64") else r32)).sum()))
# ---- timing at config 2
M, spm = 10000, 1024
segs, obs, cnt = make(M, spm, seed=3)
S, Ob, C = d(segs), d(obs), d(cnt)
S32 = d(segs[:, [1, 0, 3, 2]].astype(np.float32))
v64 = torch.empty(M * spm, dtype=torch.uint8, device="cuda"); v32 = torch.empty_like(v64)
print("old f64 %.4f ms" % t(lambda: ops.segcheck_edage_f64(S, Ob, C, clear, out=v64)))
print("old f32 %.4f ms" % t(lambda: ops.segcheck_mpnet_f32(S32, Ob, C, clear, out=v32)))
for want in (("bits64", "bits32"), ("u8_64", "u8_32"), ("bits64",), ("bits32",)):
    out = {}
    print("fused", want, "%.4f ms" % t(lambda: ops.verdict_fused(S, Ob, C, clear, want=want, out=out)))

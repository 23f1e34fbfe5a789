import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from ppnet_b200 import host, ops
from ppnet_b200.synthetic import synthetic_bank, synthetic_segments
M, SPM, R, O = 10000, 1024, 224, 50
bk = synthetic_bank(100, seed=0)
keys = ("pathpt", "segpt", "hull", "hull_cnt", "obs", "obs_cnt")
hbank = host.HostBank(*[bk[k] for k in keys], device=0)
ctx = host.HostContext(0)
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
s64 = torch.from_numpy(synthetic_segments(M, SPM, seed=1)).pin_memory().numpy()
s32 = torch.from_numpy(s64.astype(np.float32)).pin_memory().numpy()
pomax, np_, ns1 = bk["obs"].shape[1], bk["pathpt"].shape[1], bk["segpt"].shape[1]
full = dict(angle=pin([M], torch.float64), trans=pin([M, 2], torch.int32), segpt=pin([M, ns1, 2], torch.float64),
            pathpt=pin([M, np_, 2], torch.float64), obs=pin([M, O + pomax, 3], torch.float64), obs_cnt=pin([M], torch.int32),
            rand_cnt=pin([M], torch.int32), bits=pin([M, R, 7], torch.int32), tries=pin([M], torch.int32), valid=pin([M], torch.uint8))
v = [pin([M * SPM], torch.uint8) for _ in range(3)]
def run(out, checks, n=4):
    ts = []
    for i in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        host.generate_maps_host(ctx, hbank, i * M, M, 10, O, out, R, 50.0, 5.0, 1.0, 7, raster_inflate=2.24, checks=checks)
        ts.append(1e3 * (time.perf_counter() - t0))
    return ["%.2f" % t for t in ts]
chk = dict(segs_rc_f64=s64, segs_xy_f32=s32, clearance_px=4.48, verdict_f64=v[0], verdict_f32=v[1], verdict_dda=v[2])
print("H2D only (segs up, verdicts down)   ", run(dict(valid=full["valid"]), chk))
print("D2H only (labels/obs/bits down)     ", run(full, None))
print("both                                ", run(full, chk))
chk64 = dict(segs_rc_f64=s64, clearance_px=4.48, verdict_f64=v[0])
print("f64 segs only up + all down         ", run(full, chk64))
# raw copy rates
a = torch.from_numpy(s64); d = torch.empty_like(a, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(a, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("raw H2D %.1f GB/s" % (a.numel() * 8 / (t1 - t0) / 1e9))
b = torch.empty(a.shape, dtype=a.dtype).pin_memory()
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); b.copy_(d, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("raw D2H %.1f GB/s" % (a.numel() * 8 / (t1 - t0) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(a, non_blocking=True)
with torch.cuda.stream(s2): b.copy_(d, non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print("bidirectional: %.1f GB/s total" % (2 * a.numel() * 8 / (t1 - t0) / 1e9))
g_mean, g_std, g_w = ops.gmm_params(7, 10, 2, 70.0, 5.0, device="cuda")
gm, gs, gw = g_mean.cpu().numpy(), g_std.cpu().numpy(), g_w.cpu().numpy()
hg = pin([M * 1000, 2], torch.float32)
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.gmm_sample(7, i * M * 1000, M * 1000, gm, gs, gw, out=hg)
    t1 = time.perf_counter()
    host.generate_maps_host(ctx, hbank, i * M, M, 10, O, full, R, 50.0, 5.0, 1.0, 7, raster_inflate=2.24, checks=chk)
    t2 = time.perf_counter()
    print("gmm host %.2f ms, fused after gmm %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))

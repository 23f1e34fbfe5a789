import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from ppnet_b200 import ops
from ppnet_b200.synthetic import synthetic_bank, synthetic_segments
dev = torch.device("cuda", 0)
M, SPM, R, O = 10000, 1024, 224, 50
bk = synthetic_bank(100, seed=0)
bank = ops.PathBank(*[torch.from_numpy(bk[k]).to(dev) for k in ("pathpt", "segpt", "hull", "hull_cnt", "obs", "obs_cnt")])
segs32 = torch.from_numpy(synthetic_segments(M, SPM, seed=100)).to(torch.float32).to(dev)
gen = ops.generate_maps(bank, 0, M, 10, O, R, 50.0, 5.0, 1.0, 123, raster_inflate=2.24)
torch.cuda.synchronize()
def t(fn, n=12):
    out = []
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(i); b.record(); torch.cuda.synchronize(); out.append(a.elapsed_time(b))
    return ["%.3f" % x for x in out]
print("dda same maps      ", t(lambda i: ops.dda_gridcheck(gen.bits, R, segs32, want_first=False)))
def regen(i):
    ops.generate_maps(bank, i * M, M, 10, O, R, 50.0, 5.0, 1.0, 123, out=gen, raster_inflate=2.24)
for i in range(6):
    regen(i + 1); torch.cuda.synchronize()
    print("maps from %d:" % ((i + 1) * M), t(lambda j: ops.dda_gridcheck(gen.bits, R, segs32, want_first=False), 3),
          "occupancy %.3f" % (sum(bin(int(x) & 0xffffffff).count("1") for x in gen.bits[:50].flatten().tolist()) / (50 * R * R)))
v, f = ops.dda_gridcheck(gen.bits, R, segs32)
print("blocked frac", v.float().mean().item(), "mean first", f[f >= 0].float().mean().item(), "max first", f.max().item())
import subprocess, time
if len(sys.argv) > 1:
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", sys.argv[1]], stdout=subprocess.DEVNULL)
    time.sleep(1.0)
    print("with nvidia-smi -lms", sys.argv[1])
    for rep in range(3):
        print("dda", t(lambda i: ops.dda_gridcheck(gen.bits, R, segs32, want_first=False), 10))
        g_mean, g_std, g_w = ops.gmm_params(1, 10, 2, 70.0, 5.0, device=dev)
        out = torch.empty([10_000_000, 2], dtype=torch.float32, device=dev)
        print("gmm", t(lambda i: ops.gmm_sample(1, 0, 10_000_000, g_mean, g_std, g_w, out=out), 10))
    p.terminate()

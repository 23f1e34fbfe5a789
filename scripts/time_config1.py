"""Config 1 through the mirrored class API (MapGenerate(path_num=10, ..., obstacles_num=20, clearance=3).generate(100)) and
the 1 M-map case through MapGenerate.generate: wall times + a cProfile of the host side."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.getcwd())
import torch
from ppnet_b200 import edage
from ppnet_b200.edage.MapGenerate import MapGenerate
os.chdir("/tmp")
edage.seed(0)
def once(write):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mg = MapGenerate(path_num=10, resolution=224, map_size=50, obstacles_num=20, clearance=3)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    mg.generate(map_num=100, folder_path="/tmp/ppnet_cfg1", round_index=1, write_problems=write)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return 1e3 * (t1 - t0), 1e3 * (t2 - t1), mg
for _ in range(3): once(False)
r = [once(False)[:2] for _ in range(10)]
print("config 1 (class API): PathGroup %.2f ms, generate(100 maps) %.2f ms" % (sum(a for a, _ in r) / 10, sum(b for _, b in r) / 10))
pr = cProfile.Profile(); pr.enable()
for _ in range(5): once(False)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:3800])
# 1 M maps through the class API (labels only): is the host the limit?
mg = once(False)[2]
torch.cuda.synchronize(); t0 = time.perf_counter()
mg.MapLabel = []
mg.generate(map_num=200000, folder_path="/tmp/ppnet_cfg1", round_index=2, write_problems=False)
torch.cuda.synchronize(); t1 = time.perf_counter()
print("200 k maps through MapGenerate.generate: %.2f s (%d labels) -> %.2e maps/s" % (t1 - t0, len(mg.MapLabel), len(mg.MapLabel) / (t1 - t0)))

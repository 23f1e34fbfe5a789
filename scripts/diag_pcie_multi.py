"""Aggregate host<->device bandwidth of the box with 1, 2, 4, 8 GPUs copying at once (raw pinned copies, torch only; run
under torch.distributed.run with 8 ranks).  This is the ceiling of bench.py's e2e at N GPUs: every rank moves 328 MB up and
326 MB down per pass through the same host memory system."""
import json, os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = 1 << 20
tot, chunk = 328 * MB, 41 * MB
h_in = torch.empty(tot, dtype=torch.uint8).pin_memory(); h_out = torch.empty(tot, dtype=torch.uint8).pin_memory()
d_in = torch.empty(tot, dtype=torch.uint8, device="cuda"); d_out = torch.empty(tot, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(tot // chunk):
            sl = slice(i * chunk, (i + 1) * chunk)
            if up:
                with torch.cuda.stream(s1): d_in[sl].copy_(h_in[sl], non_blocking=True)
            if down:
                with torch.cuda.stream(s2): h_out[sl].copy_(d_out[sl], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
out = {}
for n in (1, 2, 4, 8):
    if n > world: break
    row = {}
    for name, up, down in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        dist.barrier(); torch.cuda.synchronize()
        t = run(up, down) if rank < n else 0.0
        tt = torch.tensor([t], device="cuda", dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        row[name + "_ms_per_328MB"] = float(tt.item()) * 1e3
        row[name + "_aggregate_GBs"] = n * tot * (2 if (up and down) else 1) / float(tt.item()) / 1e9
    out["%d_gpus" % n] = row
if rank == 0: print(json.dumps(out, indent=1))
dist.destroy_process_group()

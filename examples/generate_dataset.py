"""The reference's dataset-generation entry point (EDaGe-PP/MapGenerate.py:154-176) on ppnet_b200:

    for each round: MapGenerate(path_num=10, resolution=224, map_size=50, obstacles_num=20, clearance=3)
                    .generate(map_num=100, folder_path=..., round_index=...)

Same classes, same arguments; the imports are the only change.  BASELINE.json config 1 is one such round on the
CPU (reference: 10.3 s for the 10 target paths + 33.3 s for the 100 maps on one core).

    python examples/generate_dataset.py --rounds 5 [--out /tmp/ppnet_data] [--images]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ppnet_b200 import edage  # noqa: E402
from ppnet_b200.edage.MapGenerate import MapGenerate  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--out", default=None, help="write unsolved_problems.txt (and jpgs with --images) under this folder")
    ap.add_argument("--images", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    edage.seed(args.seed)
    if args.out:
        os.makedirs(args.out, exist_ok=True)
        os.chdir(args.out)
    times = []
    labels = 0
    for round_index in range(args.rounds + 1):                     # round 0 is the warm-up (CUDA context, lazy module load)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        map_generator = MapGenerate(path_num=10, resolution=224, map_size=50, obstacles_num=20, clearance=3)
        t1 = time.perf_counter()
        map_generator.generate(map_num=100, folder_path=os.path.join(args.out or ".", str(round_index)),
                               round_index=round_index, write_problems=bool(args.out), save_images=args.images)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if args.out and args.images:                               # MapGenerate.py:174: the label file SegNet / GenNet loaders read
            torch.save(map_generator.MapLabel, '{}/data/MapLabel'.format(os.path.join(args.out, str(round_index))))
        if round_index:
            times.append((t1 - t0, t2 - t1))
            labels += len(map_generator.MapLabel)
    p = sum(t[0] for t in times) / len(times)
    m = sum(t[1] for t in times) / len(times)
    print("rounds %d: PathGroup (10 target paths) %.1f ms, generate (100 maps) %.1f ms per round; %d labels; "
          "reference on one CPU core: 10300 ms + 33300 ms" % (len(times), 1e3 * p, 1e3 * m, labels))


if __name__ == "__main__":
    main()

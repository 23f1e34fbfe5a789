"""EDaGe-PP/gerated_by_planners.py -- rasterise OMPL planner solutions into map / mask_space / mask_path so the nets
can be trained on planner data.  Same entry points; the painting runs on the GPU (ppnet_planner_masks)."""
import json
import os

import numpy as np
import torch

from .. import ops
from .Path import _dev, plot_obstacles
from .process_map import _label_colormap, add_init_end_single

PLANNERS = ['BITstar', 'ABITstar', 'InformedRRTstar', 'RRTstar']


def load_data(data_path, planner='RRTstar', subset='train'):
    """gerated_by_planners.py:23-53: solutions of `planner` that finished in < 59 s -> (envs, paths)."""
    assert os.path.exists(data_path), "path '{}' does not exist.".format(data_path)
    assert planner in set(PLANNERS)
    assert subset in {'train', 'val', 'test'}
    envs, paths = [], []
    with open(data_path, 'r', encoding='utf-8') as f:
        for line in f:
            problem = json.loads(line)
            for s in problem["Solution"]:
                if s["Planner"] == planner and s["Waypoint"] and s["Time"] < 59:
                    paths.append(s["Waypoint"])
                    envs.append([[o[:2], o[2]] for o in problem["Obstacles"]])
    if subset in ('val', 'test'):
        envs, paths = envs[5000:], paths[5000:]
    print('{} set :'.format(subset), len(envs), 'its')
    return envs, paths


def solution_masks(paths, clearance=1 / 50 * 224):
    """The two label masks of a list of solutions [[x, y], ...] -> (mask_space, mask_path) uint8 tensors [n,224,224]."""
    dev = _dev()
    if not paths:
        z = torch.zeros([0, 224, 224], dtype=torch.uint8, device=dev)
        return z, z.clone()
    wp = torch.from_numpy(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 2) for p in paths])).to(dev)
    off = torch.from_numpy(np.cumsum([0] + [len(p) for p in paths]).astype(np.int64)).to(dev)
    out_s, out_p = [], []
    for a in range(0, len(paths), 32768):                              # launch limit: 65535 solutions
        b = min(len(paths), a + 32768)
        sub_off = (off[a:b + 1] - off[a]).contiguous()
        s, p = ops.planner_masks(wp[int(off[a]):int(off[b])].contiguous(), sub_off, clearance)
        out_s.append(s)
        out_p.append(p)
    return torch.cat(out_s), torch.cat(out_p)


def generated_by_planners(data_path):
    """gerated_by_planners.py:56-161: ./data_{planner}/{map,mask_space,mask_path}/{i}.{jpg,png} for the four planners."""
    import torchvision
    from PIL import Image
    cm = _label_colormap().flatten()
    for planner in PLANNERS:
        root = './data_{}'.format(planner)
        for sub in ('map', 'mask_space', 'mask_path'):
            os.makedirs(os.path.join(root, sub), exist_ok=True)
        envs, paths = load_data(data_path, planner)
        space, pathm = solution_masks(paths)
        space, pathm = space.cpu().numpy(), pathm.cpu().numpy()
        for i, (e, p) in enumerate(zip(envs, paths)):
            image = plot_obstacles((224, 224), [[o[0][0], o[0][1], o[1]] for o in e], resolution=(224, 224))
            image = add_init_end_single(image, [p[0][1], p[0][0]], [p[-1][1], p[-1][0]])
            torchvision.utils.save_image(image, '{}/map/{}.jpg'.format(root, i))
            m = Image.fromarray((space[i] * 255).astype(np.uint8), mode='P')     # float 0/1 -> 'P' stores 255
            m.putpalette(cm)
            m.save('{}/mask_space/{}.png'.format(root, i))
            Image.fromarray(pathm[i], mode='L').save('{}/mask_path/{}.png'.format(root, i))   # float 255 -> 'L' stores 1

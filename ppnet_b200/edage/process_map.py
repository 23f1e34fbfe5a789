"""EDaGe-PP/process_map.py -- the two functions on the hot path: the segment-vs-circles verdict (:383-425) and the
init/end stamps (:119-145).  Scalar signatures as in the reference, plus batched forms."""
import numpy as np
import torch

from .. import ops
from .Path import _dev


def _obs_tensor(obs, dev):
    o = np.asarray([[float(v[0]), float(v[1]), float(v[2])] for v in obs], dtype=np.float64).reshape(-1, 3)
    omax = max(len(o), 1)
    t = torch.zeros([1, omax, 3], dtype=torch.float64, device=dev)
    if len(o):
        t[0, :len(o)] = torch.from_numpy(o).to(dev)
    return t, torch.tensor([len(o)], dtype=torch.int32, device=dev)


def collision_check_circle_edge(s, e, obs, clearance):
    """process_map.py:383-425: s, e in (row, col); obs = [[x, y, r], ...]; -> bool (True = collision)."""
    dev = _dev()
    pts = torch.tensor([[float(s[0]), float(s[1]), float(e[0]), float(e[1])]], dtype=torch.float64, device=dev)
    o, c = _obs_tensor(obs, dev)
    return bool(ops.segcheck_edage_f64(pts, o, c, float(clearance)).item())


def collision_check_path(path, obs, clearance):
    """The checker loop of extract_path_image (process_map.py:491-495) in one launch: path = [[row, col], ...]
    -> bool array, one verdict per consecutive pair."""
    dev = _dev()
    p = np.asarray(path, dtype=np.float64).reshape(-1, 2)
    if len(p) < 2:
        return np.zeros(0, dtype=bool)
    pts = torch.from_numpy(np.ascontiguousarray(np.concatenate([p[:-1], p[1:]], axis=1))).to(dev)
    o, c = _obs_tensor(obs, dev)
    return ops.segcheck_edage_f64(pts, o, c, float(clearance)).cpu().numpy().astype(bool)


def add_init_end_single(image, init, end):
    """process_map.py:119-145: 7x7 red squares at round(init), round(end) on image Tensor[3,R,R] (in place when the
    tensor already lives on the GPU; a CPU tensor is updated through a device copy)."""
    assert len(image.shape) == 3, "Image shape incorrect"
    assert init is not None, "Init is None"
    assert end is not None, "End is None"
    dev = _dev()
    img = image.to(device=dev, dtype=torch.float32)
    work = img.contiguous()[None]
    mk = lambda p: torch.tensor([[float(p[0]), float(p[1])]], dtype=torch.float64, device=dev)
    ops.add_init_end(work, mk(init), mk(end))
    if work.data_ptr() != image.data_ptr():
        image.copy_(work[0].to(device=image.device, dtype=image.dtype))
    return image


# ---------------------------------------------------------------------------------------------------------------
# "next" rows: label-mask rasterisers and post-hoc path extraction (process_map.py:148-191, 293-365)
# ---------------------------------------------------------------------------------------------------------------
NUM_PER_FOLDER = 400


def _label_colormap(n=256):
    """The VOC-style label colormap imgviz.label_colormap() returns (bit-interleaved label index)."""
    cm = np.zeros([n, 3], dtype=np.uint8)
    for i in range(n):
        c, r, g, b = i, 0, 0, 0
        for j in range(8):
            r |= ((c >> 0) & 1) << (7 - j)
            g |= ((c >> 1) & 1) << (7 - j)
            b |= ((c >> 2) & 1) << (7 - j)
            c >>= 3
        cm[i] = [r, g, b]
    return cm


def gen_path_masks(path_point, resolution=224):
    """generate_gen_path without the file I/O: path_point [n][Np][2] (row, col) -> uint8 tensor [n,R,R] on the GPU,
    255 at every 5th label point (strictly inside the image)."""
    dev = _dev()
    pp = torch.from_numpy(np.ascontiguousarray(np.asarray(path_point, dtype=np.float64))).to(dev)
    return ops.path_mask(pp, resolution)


def generate_gen_path(path_point, folder_index, root: str = './'):
    """process_map.py:148-163: one 'L' PNG per map, {root}/{folder_index*NUM_PER_FOLDER + i}.png.  The reference's
    float image goes through ToPILImage().convert('L'), which stores painted pixels as 1 -- kept."""
    import os
    from PIL import Image
    os.makedirs(root, exist_ok=True)
    masks = (gen_path_masks(path_point) != 0).to(torch.uint8).cpu().numpy()
    for i, m in enumerate(masks):
        Image.fromarray(m, mode='L').save('{}/{}.png'.format(root, int(folder_index) * NUM_PER_FOLDER + i))


def seg_space_masks(spaces, n_maps, rotation, translation, resolution=224):
    """generate_seg_space without the file I/O: spaces = corridor masks (PIL / array / tensor, one per target path),
    map i uses spaces[int(i / (n_maps / len(spaces)))], rotated by -rotation[i] and translated by translation[i]
    (torchvision semantics), thresholded at 0.5 -> uint8 tensor [n,R,R] in {0, 1}."""
    dev = _dev()
    src = []
    for sp in spaces:
        a = np.asarray(sp)
        if a.ndim == 3:
            a = a[..., 0] if a.shape[-1] in (1, 3, 4) else a[0]
        if a.dtype != np.uint8:
            a = (np.asarray(a, dtype=np.float64) * 255).round().astype(np.uint8)
        src.append(a)
    src = torch.from_numpy(np.ascontiguousarray(np.stack(src))).to(dev)
    idx = torch.tensor([int(i / (n_maps / len(spaces))) for i in range(n_maps)], device=dev)
    ang = torch.tensor([-float(np.reshape(r, -1)[0]) for r in rotation], dtype=torch.float64, device=dev)
    tr = torch.tensor([[float(t[0]), float(t[1])] for t in translation], dtype=torch.float64, device=dev)
    placed = ops.mask_rigid(src[idx].contiguous(), ang, tr, resolution)
    return (placed > 127).to(torch.uint8)            # ToTensor scales to [0, 1]; cv2.threshold(0.5)


def generate_seg_space(spaces, path_point, rotation, translation, folder_index, root: str = './'):
    """process_map.py:166-191: one palette PNG per map (index 255 = corridor, as the reference's float -> 'P' conversion
    stores it)."""
    import os
    from PIL import Image
    os.makedirs(root, exist_ok=True)
    masks = seg_space_masks(spaces, len(path_point), rotation, translation).cpu().numpy() * 255
    cm = _label_colormap().flatten()
    for i, m in enumerate(masks):
        img = Image.fromarray(m.astype(np.uint8), mode='P')
        img.putpalette(cm)
        img.save('{}/{}.png'.format(root, int(folder_index) * NUM_PER_FOLDER + i))


def extract_path(mask, init_state, end_state, down_sample_rate=8, max_len=4096):
    """process_map.py:293-365: mask = PIL heat-map; greedy walk from init_state to end_state on the bilinearly
    down-sampled mask.  -> (True, Tensor[L, 2]) or (False, None).  The reference's 1 s timeout is a step budget here."""
    from PIL import Image
    dev = _dev()
    small = mask.resize((int(mask.size[0] / down_sample_rate), int(mask.size[1] / down_sample_rate)), Image.BILINEAR)
    a = np.asarray(small)
    a = (a.astype(np.float32) / 255.0) if a.dtype == np.uint8 else a.astype(np.float32)       # ToTensor
    if a.ndim == 3:
        a = a[..., 0]
    m = torch.from_numpy(np.ascontiguousarray(a))[None].to(dev)
    mk = lambda p: torch.tensor([[float(p[0]), float(p[1])]], dtype=torch.float64, device=dev)
    out, ln, ok = ops.extract_path(m, mk(init_state), mk(end_state), float(down_sample_rate), max_len=max_len)
    if not bool(ok.item()):
        return False, None
    return True, out[0, :int(ln.item())].cpu()


def extract_path_batch(masks, init_states, end_states, down_sample_rate=8, max_len=4096):
    """extract_path over many heat-maps of one size in ONE launch: masks = list of PIL images, init / end = [n][2].
    -> list of (ok, Tensor[L, 2] or None), entry i what `extract_path(masks[i], ...)` returns."""
    from PIL import Image
    dev = _dev()
    n = len(masks)
    if n == 0:
        return []
    small = []
    for mask in masks:
        s = mask.resize((int(mask.size[0] / down_sample_rate), int(mask.size[1] / down_sample_rate)), Image.BILINEAR)
        a = np.asarray(s)
        a = (a.astype(np.float32) / 255.0) if a.dtype == np.uint8 else a.astype(np.float32)       # ToTensor
        small.append(a[..., 0] if a.ndim == 3 else a)
    if any(a.shape != small[0].shape for a in small):
        raise ops.PPNetError("extract_path_batch: heat-maps of different sizes; call extract_path per image")
    m = torch.from_numpy(np.ascontiguousarray(np.stack(small))).to(dev)
    mk = lambda ps: torch.tensor([[float(p[0]), float(p[1])] for p in ps], dtype=torch.float64, device=dev)
    out, ln, ok = ops.extract_path(m, mk(init_states), mk(end_states), float(down_sample_rate), max_len=max_len)
    out, ln, ok = out.cpu(), ln.cpu().numpy(), ok.cpu().numpy()
    return [(True, out[i, :int(ln[i])]) if ok[i] else (False, None) for i in range(n)]


def read_folder(root: str = './', is_read_path: bool = True):
    """process_map.py:75-102: ({root}/*.jpg|png sorted by their numeric stem, the `MapLabel` list torch.save'd under
    {root}/data, the corridor images under {root}/data as arrays).  A folder that does not hold NUM_PER_FOLDER images is
    returned unsorted, as the reference does."""
    import os
    from PIL import Image
    spaces_obs = []
    supported = [".jpg", ".JPG", ".png", ".PNG"]
    labels = torch.load(r'{}/data/MapLabel'.format(root), weights_only=False)
    if is_read_path:
        spaces_folder = r'{}/data'.format(root)
        images = [i for i in os.listdir(spaces_folder) if os.path.splitext(i)[-1] in supported]
        for i in list(images):
            images[int(i[0:-4])] = i
        for im in images:
            spaces_obs.append(np.asarray(Image.open(os.path.join(spaces_folder, im))).copy())
    images = [i for i in os.listdir(root) if os.path.splitext(i)[-1] in supported]
    if len(images) != NUM_PER_FOLDER:
        return [os.path.join(root, i) for i in images], labels, spaces_obs
    for i in list(images):
        images[int(i[0:-4])] = i
    images = [os.path.join(root, i) for i in images]
    return images, labels[0:len(images)], spaces_obs


def extract_path_image(mask_root, origin_data_root, result_root, unsolved_problems_txt, clearance, plot=None):
    """process_map.py:452-506, the caller of A11 on network output: for every heat-map {mask_root}/{index}.png with a
    matching entry in `unsolved_problems_txt`, extract the waypoints between the map's first and last segment point
    (`down_sample_rate` 2), reject the path if any edge collides with the problem's obstacles, else append the problem
    with its `Length` and `Waypoint` to {result_root}/solved_problems.txt -- the same lines in the same order.  All
    extractions run in one launch per mask size and all edges of all paths in ONE A11 launch (CSR by path).  The
    reference draws every solution with matplotlib (`plot_solution`); pass `plot=callable(obstacles, path, index)`
    to do the same.  -> list of failed indices (the reference prints it)."""
    import json
    import os
    from PIL import Image
    from scipy.spatial import distance
    with open(unsolved_problems_txt, 'r', encoding='utf-8') as f:
        unsolved_problems = [json.loads(line) for line in f.readlines()]
    solved_problems_txt = result_root + "/solved_problems.txt"
    if not os.path.exists(result_root):
        os.mkdir(result_root)
    supported = [".png", ".PNG"]
    masks = sorted(os.path.join(mask_root, i) for i in os.listdir(mask_root) if os.path.splitext(i)[-1] in supported)
    by_index = {}
    for p in unsolved_problems:                     # the reference keeps the FIRST problem with a matching index
        by_index.setdefault(p["Index"], p)
    folders = {}
    jobs = []                                       # (index string, problem, PIL mask, init, end)
    for m in masks:
        index = m.split('/')[-1].split('.')[0]
        problem = by_index.get(int(index))
        if problem is None:
            print("No matched problem(Index:{})".format(index))
            continue
        folder_index, img_index = int(int(index) / NUM_PER_FOLDER), int(int(index) % NUM_PER_FOLDER)
        if folder_index not in folders:
            folders[folder_index] = read_folder(os.path.join(origin_data_root, str(folder_index)), is_read_path=True)
        images, labels, _ = folders[folder_index]
        if len(images) != NUM_PER_FOLDER or len(labels) != NUM_PER_FOLDER:
            continue
        jobs.append((index, problem, Image.open(m), labels[img_index][3][0], labels[img_index][3][10]))
    # one extraction launch per mask size
    results = [None] * len(jobs)
    by_size = {}
    for k, j in enumerate(jobs):
        by_size.setdefault(j[2].size, []).append(k)
    for ks in by_size.values():
        got = extract_path_batch([jobs[k][2] for k in ks], [jobs[k][3] for k in ks], [jobs[k][4] for k in ks], down_sample_rate=2)
        for k, g in zip(ks, got):
            results[k] = g
    # every edge of every extracted path against its problem's obstacles: one A11 launch, paths as CSR rows
    dev = _dev()
    live = [k for k, r in enumerate(results) if r[0]]
    blocked = {}
    if live:
        omax = max(1, max(len(jobs[k][1]["Obstacles"]) for k in live))
        obs = torch.zeros([len(live), omax, 3], dtype=torch.float64)
        cnt = torch.zeros([len(live)], dtype=torch.int32)
        edges, off = [], [0]
        for r, k in enumerate(live):
            o = jobs[k][1]["Obstacles"]
            if len(o):
                obs[r, :len(o)] = torch.tensor([[float(v[0]), float(v[1]), float(v[2])] for v in o], dtype=torch.float64)
            cnt[r] = len(o)
            p = results[k][1].to(torch.float64)
            edges.append(torch.cat([p[:-1], p[1:]], dim=1))
            off.append(off[-1] + len(p) - 1)
        v = ops.segcheck_edage_f64(torch.cat(edges).contiguous().to(dev), obs.to(dev), cnt.to(dev), float(clearance),
                                   seg_off=torch.tensor(off, dtype=torch.int64, device=dev)).cpu().numpy()
        for r, k in enumerate(live):
            blocked[k] = bool(v[off[r]:off[r + 1]].any())
    failure_cases = []
    for k, (index, problem, _, _, _) in enumerate(jobs):
        rst, path = results[k]
        if not rst:
            print("Extract failed:", index)
            failure_cases.append(index)
            continue
        if blocked[k]:
            failure_cases.append(index)
            print("Collision:", index)
            continue
        print("Planning succeed:", index)
        length = sum([distance.euclidean(path[i], path[i + 1]) for i in range(len(path) - 1)])
        problem["Length"] = length
        problem["Waypoint"] = [list(p) for p in list(path.numpy())]
        with open(solved_problems_txt, "a") as f:
            f.write(json.dumps(problem) + "\n")
        print('cost:', length)
        if plot is not None:
            plot(problem["Obstacles"], path, index)
    print(len(failure_cases), failure_cases)
    return failure_cases

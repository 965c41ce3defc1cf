"""Ultra-resolution patch-grid sampler: the reference's sample_ultra_res.py entry points (scope rows a9-a15 of SURVEY.md
section 8) with the same names, positional signatures and bit-exact integer geometry, re-hosted for one process per GPU.

Reference                                   | here
--------------------------------------------+---------------------------------------------------------------------------
mp.Process per GPU + mp.Queue + Manager dict | SPMD under torchrun: every rank computes the same static, dependency-driven
(sample_ultra_res.py:213-261), whole patches | schedule (`build_schedule`); a patch-stage runs on one rank; only the overlap
pickled through the CPU, busy-wait re-queue  | border strips a dependent patch needs travel, GPU to GPU, as one grouped
(:141-143), model reload per stage (:79)     | NCCL send/recv batch per round; models of all stages stay resident.
inpaint canvas on the CPU (:149-170)         | kd_border_pack kernel reading resident patches or received strips.
torch.roll of the whole image per patch      | padded index gather of the 1024^2 window only (`get_cond_images`).
(:358-398)                                   |

Results are independent of the number of GPUs: every patch draws counter-based noise keyed by (magnification, patch index)
and every kernel's reduction order is independent of the batch a patch is in.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn.functional as F

PATCH_SIZE = 1024                              # sample_ultra_res.py:31
PATCH_SIZES = {1: 64, 2: 256, 3: 1024}         # sample_ultra_res.py:32
MAG_LEVEL_SIZES = [40000, 6500, 1024]          # ultra_res_patient_dataset.py:18
AIRS_MAG_LEVEL_SIZES = [10000, 3328, 1024]     # ultra_res_airs.py:23
MAX_BATCH = {1: 16, 2: 8, 3: 2}                # patches of one round batched per rank, per stage (64^2 / 256^2 / 1024^2)


# ------------------------------------------------------------------------------------------------ geometry (a12-a14)
def get_patch_width(args, mag_level):
    """sample_ultra_res.py:273-280 (Python float -> int truncation kept)."""
    sizes = AIRS_MAG_LEVEL_SIZES if getattr(args, "version", None) == "airs" else MAG_LEVEL_SIZES
    return int(sizes[mag_level] * PATCH_SIZE / sizes[mag_level - 1])


def _center_crop_index(size_in, size_out):
    """torchvision CenterCrop along one axis as (source index per output position, validity): pads with zeros when the
    input is smaller ((out-in)//2 before, (out-in+1)//2 after), crops from int(round((in-out)/2.0)) otherwise."""
    if size_in < size_out:
        pad_lo = (size_out - size_in) // 2
        padded = size_in + pad_lo + (size_out - size_in + 1) // 2
        top = int(round((padded - size_out) / 2.0))
        src = torch.arange(size_out) + top - pad_lo
    else:
        top = int(round((size_in - size_out) / 2.0))
        src = torch.arange(size_out) + top
    valid = (src >= 0) & (src < size_in)
    return src.clamp(0, size_in - 1), valid


def _rgb_to_hsv(img):
    """skimage.color.rgb2hsv on an (H,W,3) float array (skimage is not installed here: restated, unpinned)."""
    import numpy as np

    out = np.empty_like(img)
    v = img.max(-1)
    delta = img.max(-1) - img.min(-1)
    old = np.seterr(invalid="ignore", divide="ignore")
    s = delta / v
    s[delta == 0.0] = 0.0
    h = np.zeros_like(v)
    r, g, b = img[..., 0], img[..., 1], img[..., 2]
    idx = r == v
    h[idx] = (g[idx] - b[idx]) / delta[idx]
    idx = g == v
    h[idx] = 2.0 + (b[idx] - r[idx]) / delta[idx]
    idx = b == v
    h[idx] = 4.0 + (r[idx] - g[idx]) / delta[idx]
    h = (h / 6.0) % 1.0
    h[delta == 0.0] = 0.0
    np.seterr(**old)
    out[..., 0], out[..., 1], out[..., 2] = h, s, v
    out[np.isnan(out)] = 0
    return out


def get_cond_images(args, zoomed_image, mag_level):
    """sample_ultra_res.py:304-400.  Returns (cond_images [P, 3|6, 1024, 1024], patch_pos, num_patches_width).

    Per patch the reference rolls the whole zoomed image so the patch centre lands on the image centre, overwrites the
    wrapped-around rows / columns with the fill colour and centre-crops 1024^2; here only the 1024^2 window is gathered
    with index vectors that encode roll, fill and crop (bit-identical output, including the shift == 0 quirk of
    :380-388 where the `else` branch fills the whole axis)."""
    patch_width = get_patch_width(args, mag_level)
    patch_dist = int(patch_width * (1 - args.overlap))
    W = zoomed_image.shape[3]
    num_patches_width = 1 + math.ceil((W - patch_width) / patch_dist)
    airs = getattr(args, "version", None) == "airs"
    if airs:
        num_patches_width = max(1, num_patches_width - 1)

    if mag_level == 2:
        import cv2
        import numpy as np

        img_np = zoomed_image[0].permute(1, 2, 0).cpu().numpy()
        hsv = _rgb_to_hsv(img_np)
        keep = hsv[:, :, 2] > 0.1 if airs else np.logical_and(hsv[:, :, 0] > 0.5, hsv[:, :, 1] > 0.02)
        keep = cv2.erode(keep.astype(np.uint8), np.ones((5, 5), np.uint8), iterations=1)
        keep = cv2.dilate(keep.astype(np.uint8), np.ones((51, 51), np.uint8), iterations=1)
        patch_pos = []
        for i in range(num_patches_width):
            for j in range(num_patches_width):
                y, x = i * patch_dist, j * patch_dist
                if np.any(keep[y:y + patch_width, x:x + patch_width] > 0.5):
                    patch_pos.append((i, j))
    else:
        patch_pos = [(i, j) for i in range(num_patches_width) for j in range(num_patches_width)]

    fill = 0.0 if airs else 0.95
    src_img = zoomed_image[0]
    dev = src_img.device
    crop_src, crop_valid = _center_crop_index(W, PATCH_SIZE)  # positions in the shifted image

    def axis_index(shift):
        # shifted[p] = zoomed[(p - shift) mod W]; rows [0, shift) are filled when shift > 0, rows [W + shift, W) otherwise
        src = (crop_src - shift) % W
        filled = (crop_src < shift) if shift > 0 else (crop_src >= W + shift)
        return src.to(dev), filled.to(dev)

    cond_images = []
    cvalid = crop_valid.to(dev)
    for i, j in patch_pos:
        y, x = i * patch_dist, j * patch_dist
        shift_y = W // 2 - (y + patch_width // 2)
        shift_x = W // 2 - (x + patch_width // 2)
        iy, fy = axis_index(shift_y)
        ix, fx = axis_index(shift_x)
        cond = src_img.index_select(1, iy).index_select(2, ix)
        filled = fy[:, None] | fx[None, :]
        cond = torch.where(filled[None], torch.full((), fill, dtype=cond.dtype, device=dev), cond)
        pad = ~(cvalid[:, None] & cvalid[None, :])
        cond = torch.where(pad[None], torch.zeros((), dtype=cond.dtype, device=dev), cond)
        if getattr(args, "version", None) == "v2":
            top = int(round((PATCH_SIZE - patch_width) / 2.0))
            center = cond[:, top:top + patch_width, top:top + patch_width]
            center = F.interpolate(center.unsqueeze(0), PATCH_SIZE, mode="nearest").squeeze(0)
            cond = torch.cat((cond, center), 0)
        cond_images.append(cond)
    return torch.stack(cond_images), patch_pos, num_patches_width


def get_next_patches(patches, orientation):
    """sample_ultra_res.py:403-412 (roots of the dependency DAG), with a set instead of O(P^2) list scans."""
    present = set(map(tuple, patches))
    processed, waiting = [], []
    for i, j in patches:
        if (i - 1, j) not in present and (i, j + orientation) not in present and (i - 1, j + orientation) not in present:
            processed.append((i, j))
        else:
            waiting.append((i, j))
    return processed, waiting


def choose_orientation(patch_pos):
    """sample_ultra_res.py:423-426 (strict '>')."""
    left = len(get_next_patches(patch_pos, -1)[0])
    right = len(get_next_patches(patch_pos, 1)[0])
    return -1 if left > right else 1


# ------------------------------------------------------------------------------------------------ schedule (a9 / a10, section 8e)
def neighbours(pos, orientation):
    i, j = pos
    return dict(above=(i - 1, j), side=(i, j + orientation), corner=(i - 1, j + orientation))


def dependents(pos, orientation):
    """Patches that use `pos` as their above / side / corner neighbour, and which strip of `pos` each needs."""
    i, j = pos
    return dict(above=(i + 1, j), side=(i, j - orientation), corner=(i + 1, j - orientation))


class Schedule:
    def __init__(self, rounds, owner):
        self.rounds, self.owner = rounds, owner  # rounds: list of {rank: [patch index, ...]}


def build_schedule(patch_pos, orientation, world, max_batch=1):
    """Deterministic dependency-driven list schedule (identical on every rank).  A patch is ready once its (up to three)
    neighbours that exist in the grid are done (sample_ultra_res.py:99-107).  Each round takes up to world * max_batch
    ready patches, longest-remaining-dependency-chain first, spread evenly; a patch prefers the rank that produced its
    `above` neighbour (that strip is the largest and then stays local)."""
    patch_pos = [tuple(p) for p in patch_pos]
    index = {p: k for k, p in enumerate(patch_pos)}
    deps = [[index[nb] for nb in neighbours(p, orientation).values() if nb in index] for p in patch_pos]
    users = [[index[q] for q in dependents(p, orientation).values() if q in index] for p in patch_pos]
    # priority = length of the longest chain of dependents hanging off a patch
    order = sorted(range(len(patch_pos)), key=lambda k: (-patch_pos[k][0], orientation * patch_pos[k][1]))
    prio = [0] * len(patch_pos)
    for k in order:  # bottom rows first; within a row, patches further along the dependency direction first
        prio[k] = 1 + max((prio[u] for u in users[k]), default=0)
    done, remaining, rounds, owner = set(), set(range(len(patch_pos))), [], {}
    while remaining:
        ready = sorted((k for k in remaining if all(d in done for d in deps[k])), key=lambda k: (-prio[k], patch_pos[k]))
        assert ready, "dependency cycle in the patch grid"
        chosen = ready[: world * max_batch]
        cap = -(-len(chosen) // world)
        load = [0] * world
        rnd = {}
        for k in chosen:
            i, j = patch_pos[k]
            above = index.get((i - 1, j))
            pref = owner[above] if above is not None else (j % world)
            r = pref if load[pref] < cap else min(range(world), key=lambda q: (load[q], q))
            load[r] += 1
            owner[k] = r
            rnd.setdefault(r, []).append(k)
        rounds.append(rnd)
        done.update(chosen)
        remaining.difference_update(chosen)
    return Schedule(rounds, owner)


# ------------------------------------------------------------------------------------------------ distributed plumbing
class PatchSet(list):
    """Per-rank view of a stage's outputs: entry k is a (1,3,S,S) tensor on this rank or None; `owner[k]` says where it
    lives (None = replicated on every rank)."""

    def __init__(self, items, owner=None):
        super().__init__(items)
        self.owner = owner


DISABLE_DIST = False  # tests: run the single-rank schedule inside an initialised process group


def _dist():
    import torch.distributed as dist

    if not DISABLE_DIST and dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def _exchange(transfers, rank, device):
    """transfers: list of (src_rank, dst_rank, key, tensor_or_None, shape).  One grouped batch of P2P ops; returns
    {key: tensor} of what this rank received."""
    dist, _, world = _dist()
    got, ops_, bufs = {}, [], []
    for src, dst, key, tensor, shape in transfers:
        if src == dst:
            if rank == dst:
                got[key] = tensor
            continue
        if rank == src:
            ops_.append(dist.P2POp(dist.isend, tensor.contiguous(), dst))
        elif rank == dst:
            buf = torch.empty(shape, device=device, dtype=torch.float32)
            bufs.append((key, buf))
            ops_.append(dist.P2POp(dist.irecv, buf, src))
    if ops_:
        for req in dist.batch_isend_irecv(ops_):
            req.wait()
    for key, buf in bufs:
        got[key] = buf
    return got


# ------------------------------------------------------------------------------------------------ hooks (tests / bench)
def default_model_provider(mag_level, unet_number, device, args):
    """load_model of the reference (sample_ultra_res.py:36-65): factory by --version, checkpoint from args, strict load
    with a restore_parts fallback.  Models are cached per (mag, stage): they stay resident instead of being reloaded."""
    from .factories import init_imagen_ultra_res
    from .trainer import __version__, restore_parts

    version = getattr(args, "version", None) or ""
    imagen = init_imagen_ultra_res(mag_level, unet_number, device=device, version=version)
    path = vars(args)[f"unet{unet_number}_mag{mag_level}"]
    loaded = torch.load(path, map_location="cpu")
    if str(loaded.get("version", __version__)) != __version__:
        print(f'loading saved imagen at version {loaded["version"]}, but current package version is {__version__}')
    try:
        imagen.load_state_dict(loaded["model"], strict=True)
    except RuntimeError:
        print("Failed loading state dict. Trying partial load")
        imagen.load_state_dict(restore_parts(imagen.state_dict(), loaded["model"]))
    return imagen


def default_canvas(S, overlap_pos, orientation, above, side, corner, device):
    from . import ops

    return ops.border_pack(S, overlap_pos, orientation, above, side, corner, device)


MODEL_PROVIDER = default_model_provider
CANVAS_FN = default_canvas
_MODEL_CACHE = {}


def load_model(mag_level, unet_number, device, args):
    key = (mag_level, unet_number, str(device), id(MODEL_PROVIDER))
    if key not in _MODEL_CACHE:
        _MODEL_CACHE[key] = MODEL_PROVIDER(mag_level, unet_number, device, args)
    return _MODEL_CACHE[key]


def _device_for(args, rank):
    dev = getattr(args, "device", None)
    if dev is not None:
        return torch.device(dev)
    return torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))


# ------------------------------------------------------------------------------------------------ a9 / a10: one stage over the grid
def _strip_of(patch, kind, S, ov, orientation):
    """(view, channel stride, row stride) of the strip of a full resident patch [1,3,S,S] a dependent needs."""
    p = patch.reshape(3, S, S)
    x0 = S - ov if orientation == -1 else 0
    if kind == "above":
        return p[:, S - ov:, :], S * S, S
    if kind == "side":
        return p[:, :, x0:x0 + ov], S * S, S
    return p[:, S - ov:, x0:x0 + ov], S * S, S


def _strip_shape(kind, S, ov):
    return {"above": (3, ov, S), "side": (3, S, ov), "corner": (3, ov, ov)}[kind]


def generate_image_with_unet(mag_level, unet_number, args, lowres_image, cond_image, patch_pos, overlap, orientation, num_patches_width):
    """sample_ultra_res.py:213-261 (dispatcher) + :75-210 (worker), SPMD.  Returns a PatchSet of (1,3,S,S) tensors."""
    dist, rank, world = _dist()
    device = _device_for(args, rank)
    imagen = load_model(mag_level, unet_number, device, args)
    S = PATCH_SIZES[unet_number]
    n = len(patch_pos) if patch_pos is not None else (cond_image.shape[0] if cond_image is not None else 1)
    sample_kw = dict(return_pil_images=False, start_at_unet_number=unet_number, stop_at_unet_number=unet_number,
                     inpaint_resample_times=args.inpaint_resample, use_tqdm=False, device=device)

    if patch_pos is None:  # magnification 0: plain sampling, replicated on every rank (deterministic noise)
        out = []
        for idx in range(n):
            lr = None if lowres_image is None else lowres_image[idx].to(device)
            ci = None if cond_image is None else cond_image[idx].unsqueeze(0)
            out.append(imagen.sample(batch_size=1, cond_images=ci, start_image_or_video=lr, inpaint_images=None, inpaint_masks=None,
                                     noise_key=[_noise_key(mag_level, idx)], **sample_kw))
        return PatchSet(out, owner=None)

    patch_pos = [tuple(p) for p in patch_pos]
    index = {p: k for k, p in enumerate(patch_pos)}
    sched = build_schedule(patch_pos, orientation, world, max(1, min(getattr(args, "max_batch", None) or MAX_BATCH[unet_number], n)))
    owner = sched.owner
    overlap_pos = int(overlap * S)
    patch_width = get_patch_width(args, mag_level) if cond_image is not None else 0  # only used for fallback neighbour crops
    patch_dist = int(patch_width * (1 - overlap))

    # previous-stage outputs must sit on the rank that runs the patch now
    lowres_local = {}
    if lowres_image is not None:
        prev_owner = getattr(lowres_image, "owner", None)
        if prev_owner is None or world == 1:
            lowres_local = {k: lowres_image[k] for k in range(n) if owner[k] == rank}
        else:
            s_prev = PATCH_SIZES[unet_number - 1]
            moves = [(prev_owner[k], owner[k], k, None if lowres_image[k] is None else lowres_image[k].to(device), (1, 3, s_prev, s_prev))
                     for k in range(n)]
            lowres_local = _exchange(moves, rank, device)

    results, ghosts = {}, {}
    for rnd in sched.rounds:
        mine = rnd.get(rank, [])
        if mine:
            inpaints, masks = [], []
            for k in mine:
                i, j = patch_pos[k]
                strips = {}
                for kind, nb in neighbours((i, j), orientation).items():
                    if nb in index:
                        kk = index[nb]
                        if kk in results:
                            strips[kind] = _strip_of(results[kk], kind, S, overlap_pos, orientation)
                        else:
                            g = ghosts[(kk, kind)]
                            strips[kind] = (g, g.shape[1] * g.shape[2], g.shape[2])
                    else:
                        fb = _fallback_neighbour(kind, (i, j), orientation, num_patches_width, None if cond_image is None else cond_image[k],
                                                 patch_width, patch_dist, S, device)
                        strips[kind] = None if fb is None else _strip_of(fb, kind, S, overlap_pos, orientation)
                ip, im = CANVAS_FN(S, overlap_pos, orientation, strips["above"], strips["side"], strips["corner"], device)
                inpaints.append(ip)
                masks.append(im)
            lr = None if lowres_image is None else torch.cat([lowres_local[k].to(device) for k in mine], 0)
            ci = None if cond_image is None else torch.stack([cond_image[k] for k in mine]).to(device)
            out = imagen.sample(batch_size=len(mine), cond_images=ci, start_image_or_video=lr, inpaint_images=torch.stack(inpaints),
                                inpaint_masks=torch.stack(masks), noise_key=[_noise_key(mag_level, k) for k in mine], **sample_kw)
            for b, k in enumerate(mine):
                results[k] = out[b:b + 1].contiguous()
        if world > 1:
            transfers = []
            for r in sorted(rnd):
                for k in rnd[r]:
                    for kind, q in dependents(patch_pos[k], orientation).items():
                        if q in index and owner[index[q]] != r:
                            t = _strip_of(results[k], kind, S, overlap_pos, orientation)[0] if r == rank else None
                            transfers.append((r, owner[index[q]], (k, kind), t, _strip_shape(kind, S, overlap_pos)))
            ghosts.update(_exchange(transfers, rank, device))
    return PatchSet([results.get(k) for k in range(n)], owner=dict(owner))


def _noise_key(mag_level, idx):
    mag = mag_level if isinstance(mag_level, int) else 7  # "outpaint" grid
    return (mag << 24) | idx


def _fallback_neighbour(kind, pos, orientation, num_patches_width, cond_image, patch_width, patch_dist, S, device):
    """Neighbour cell that is inside the image but was filtered out of the grid: a crop of the conditioning image,
    bilinearly upsampled to the stage resolution (sample_ultra_res.py:113-140); None at the image border (:122-123)."""
    i, j = pos
    space_above = i != 0
    space_side = (orientation == 1 and j < num_patches_width - 1) or (orientation == -1 and j > 0)
    ok = dict(above=space_above, side=space_side, corner=space_above and space_side)[kind]
    if not ok or cond_image is None:
        return None
    ty = cond_image.shape[1] // 2 - patch_width // 2
    tx = cond_image.shape[2] // 2 - patch_width // 2
    y = ty - patch_dist if kind in ("above", "corner") else ty
    x = tx + orientation * patch_dist if kind in ("side", "corner") else tx
    crop = cond_image[:3, y:y + patch_width, x:x + patch_width].unsqueeze(0).to(device).float()
    return F.interpolate(crop, size=(S, S), mode="bilinear", align_corners=False).contiguous()


# ------------------------------------------------------------------------------------------------ a11 / a15
def generate_image(mag_level, args, cond_image=None, patch_pos=None, overlap=0.25, orientation=-1, num_patches_width=1, lowres_image=None):
    """sample_ultra_res.py:264-270: stage-major cascade 64 -> 256 -> 1024 over all patches."""
    if lowres_image is None:
        lowres_image = generate_image_with_unet(mag_level, 1, args, None, cond_image, patch_pos, overlap, orientation, num_patches_width)
    medres_image = generate_image_with_unet(mag_level, 2, args, lowres_image, cond_image, patch_pos, overlap, orientation, num_patches_width)
    highres_image = generate_image_with_unet(mag_level, 3, args, medres_image, cond_image, patch_pos, overlap, orientation, num_patches_width)
    return highres_image


def gather_patches(patches, dst=0):
    """Bring every patch of a distributed PatchSet to rank `dst` (as CPU tensors, like the reference's results)."""
    dist, rank, world = _dist()
    owner = getattr(patches, "owner", None)
    if world == 1 or owner is None:
        return [p.cpu() for p in patches]
    local = next(p for p in patches if p is not None)
    moves = [(owner[k], dst, k, patches[k], tuple(local.shape)) for k in range(len(patches))]
    got = _exchange(moves, rank, local.device)
    return [got[k].cpu() for k in range(len(patches))] if rank == dst else None


def generate_high_res_image(zoomed_image, mag_level, args):
    """sample_ultra_res.py:415-448.  Returns the stitched image on rank 0 (None elsewhere)."""
    cond_images, patch_pos, num_patches_width = get_cond_images(args, zoomed_image, mag_level)
    patch_width = get_patch_width(args, mag_level)
    if getattr(args, "ignore_unet_1", False):
        top = int(round((PATCH_SIZE - patch_width) / 2.0))
        lowres_image = [c[:, top:top + patch_width, top:top + patch_width].unsqueeze(0) for c in cond_images]
    else:
        lowres_image = None
    orientation = choose_orientation(patch_pos)
    mag_images = generate_image(mag_level, args, cond_image=cond_images, patch_pos=patch_pos, overlap=args.overlap, orientation=orientation,
                                num_patches_width=num_patches_width, lowres_image=lowres_image)
    mag_images = gather_patches(mag_images)
    if mag_images is None:
        return None
    return stitch(zoomed_image, mag_images, patch_pos, num_patches_width, args.overlap)


def stitch(zoomed_image, mag_images, patch_pos, num_patches_width, overlap):
    """sample_ultra_res.py:430-446: canvas = bilinear upsample of the zoomed image; patches pasted row-major, later
    patches overwrite earlier ones, no blending."""
    patch_dist = int(PATCH_SIZE * (1 - overlap))
    full_image_width = PATCH_SIZE + (num_patches_width - 1) * patch_dist
    full_image = F.interpolate(zoomed_image.cpu().float(), size=(full_image_width, full_image_width), mode="bilinear", align_corners=False)
    for index, (i, j) in enumerate(patch_pos):
        y, x = i * patch_dist, j * patch_dist
        full_image[0, :, y:y + PATCH_SIZE, x:x + PATCH_SIZE] = mag_images[index][0].cpu()
    return full_image

"""Ultra-resolution patch-grid sampler: the reference's sample_ultra_res.py entry points (scope rows a9-a15 of SURVEY.md
section 8) with the same names, positional signatures and bit-exact integer geometry, re-hosted for one process per GPU.

Reference                                   | here
--------------------------------------------+---------------------------------------------------------------------------
mp.Process per GPU + mp.Queue + Manager dict | SPMD under torchrun: every rank computes the same static, dependency-driven
(sample_ultra_res.py:213-261), whole patches | plan (grid_plan.py); a patch-stage runs on one rank; only the overlap border
pickled through the CPU, busy-wait re-queue  | strips a dependent patch on ANOTHER GPU needs travel, as one-sided NVLink stores
(:141-143), model reload per stage (:79)     | into that GPU's CUDA-IPC mailbox (grid_exec.PeerMailbox); models stay resident.
stage-major cascade (:264-270)               | one plan over all three stages: the 64^2 / 256^2 stages of later anti-diagonals
                                             | run while earlier ones are in their 1024^2 stage (pipelined wavefront).
inpaint canvas on the CPU (:149-170)         | kd_border_pack kernel reading resident patches or mailbox slots.
torch.roll of the whole image per patch      | kd_cond_gather: the 1024^2 window only, computed on demand (`CondBank`), never
(:358-398), all windows kept in host memory  | materialised for the whole grid.
CPU paste loop (:440-446)                    | kd_canvas_fill / kd_patch_paste: owner-computes stitch into every rank's canvas.

Results are independent of the number of GPUs and of the plan: every patch draws counter-based noise keyed by (run seed,
magnification, patch index) and every kernel's reduction order is independent of the batch a patch is in.
"""
from __future__ import annotations

import math
import os
import time

import torch
import torch.nn.functional as F

from . import grid_plan
from .grid_plan import dependents, neighbours  # noqa: F401  (re-exported: tests and callers use grid.neighbours)

PATCH_SIZE = 1024                              # sample_ultra_res.py:31
PATCH_SIZES = {1: 64, 2: 256, 3: 1024}         # sample_ultra_res.py:32
MAG_LEVEL_SIZES = [40000, 6500, 1024]          # ultra_res_patient_dataset.py:18
AIRS_MAG_LEVEL_SIZES = [10000, 3328, 1024]     # ultra_res_airs.py:23
MAX_BATCH = dict(grid_plan.DEFAULT_MAX_BATCH)  # patches of one stage batched per rank (64^2 / 256^2 / 1024^2)
LAST_RUN = {}                                  # diagnostics of the most recent plan execution (bench.py reports them)


# ------------------------------------------------------------------------------------------------ geometry (a12-a14)
def get_patch_width(args, mag_level):
    """sample_ultra_res.py:273-280 (Python float -> int truncation kept)."""
    sizes = AIRS_MAG_LEVEL_SIZES if getattr(args, "version", None) == "airs" else MAG_LEVEL_SIZES
    return int(sizes[mag_level] * PATCH_SIZE / sizes[mag_level - 1])


def _center_crop_offset(size_in, size_out):
    """torchvision CenterCrop along one axis: output position o reads input position o + offset; positions outside
    [0, size_in) are the zero padding ((out-in)//2 before, (out-in+1)//2 after) applied when the input is smaller."""
    if size_in < size_out:
        pad_lo = (size_out - size_in) // 2
        padded = size_in + pad_lo + (size_out - size_in + 1) // 2
        return int(round((padded - size_out) / 2.0)) - pad_lo
    return int(round((size_in - size_out) / 2.0))


def _rgb_to_hsv(img):
    """skimage.color.rgb2hsv on an (H,W,3) float array (skimage is not installed here; restated from its published
    algorithm and pinned against the standard library's colorsys in tests/test_grid_cpu.py)."""
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


def tissue_patches(args, zoomed_image, patch_width, patch_dist, num_patches_width):
    """Magnification-2 background filter of get_cond_images (sample_ultra_res.py:317-352): HSV threshold, erode 5x5, dilate
    51x51, keep grid cells whose window contains any tissue pixel.  Runs on the host exactly as the reference does (numpy +
    OpenCV morphology; integer result)."""
    import cv2
    import numpy as np

    airs = getattr(args, "version", None) == "airs"
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
    return patch_pos


class CondBank:
    """The per-patch conditioning windows of get_cond_images as a lazy sequence: entry k is the (3|6, 1024, 1024) window of
    patch k, gathered from the zoomed image when asked for (kd_cond_gather on a CUDA image, index arithmetic on the host
    otherwise) -- the whole-grid tensor (5.5 GB for 21 x 21, 35 GB for the 53 x 53 grid) is never built unless
    `materialize()` is called.  Per patch the reference rolls the whole zoomed image so the patch centre lands on the image
    centre, overwrites the wrapped-around rows / columns with the fill colour and centre-crops 1024^2 (:358-395); the same
    result follows from index arithmetic on the window alone, including the shift == 0 case in which the reference's
    `img[:, shift:, :] = FILL` branch fills the whole axis."""

    def __init__(self, zoomed, shifts, fill, channels, patch_width):
        self.zoomed = zoomed[0].float().contiguous()  # [3, W, W]
        self.shifts, self.fill, self.channels, self.patch_width = shifts, float(fill), channels, patch_width
        self.W = self.zoomed.shape[-1]
        self.off = _center_crop_offset(self.W, PATCH_SIZE)
        self.center_top = int(round((PATCH_SIZE - patch_width) / 2.0))

    @property
    def shape(self):
        return torch.Size((len(self.shifts), self.channels, PATCH_SIZE, PATCH_SIZE))

    @property
    def device(self):
        return self.zoomed.device

    def __len__(self):
        return len(self.shifts)

    def to(self, device):
        device = torch.device(device)
        if device == self.zoomed.device:
            return self
        return CondBank(self.zoomed[None].to(device), self.shifts, self.fill, self.channels, self.patch_width)

    def __iter__(self):
        return (self[k] for k in range(len(self)))

    def __getitem__(self, k):
        if isinstance(k, slice):
            return torch.stack([self[i] for i in range(*k.indices(len(self)))])
        shift_y, shift_x = self.shifts[k]
        if self.zoomed.is_cuda:
            from . import ops

            out = torch.empty((self.channels, PATCH_SIZE, PATCH_SIZE), device=self.zoomed.device, dtype=torch.float32)
            return ops.cond_gather(self.zoomed, out, self.off, shift_y, shift_x, self.fill, self.patch_width, self.center_top)
        return self._host_window(shift_y, shift_x)

    def _axis(self, shift):
        W = self.W
        p = torch.arange(PATCH_SIZE) + self.off
        valid = (p >= 0) & (p < W)
        filled = (p < shift) if shift > 0 else ((p >= W + shift) if shift < 0 else torch.ones_like(valid))
        return ((p - shift) % W).clamp(0, W - 1), filled, valid

    def _host_window(self, shift_y, shift_x):
        iy, fy, vy = self._axis(shift_y)
        ix, fx, vx = self._axis(shift_x)
        cond = self.zoomed.index_select(1, iy).index_select(2, ix)
        cond = torch.where((fy[:, None] | fx[None, :])[None], torch.full((), self.fill, dtype=cond.dtype), cond)
        cond = torch.where((~(vy[:, None] & vx[None, :]))[None], torch.zeros((), dtype=cond.dtype), cond)
        if self.channels == 6:
            t, pw = self.center_top, self.patch_width
            center = F.interpolate(cond[:, t:t + pw, t:t + pw].unsqueeze(0), PATCH_SIZE, mode="nearest").squeeze(0)
            cond = torch.cat((cond, center), 0)
        return cond

    def materialize(self):
        return torch.stack([self[k] for k in range(len(self))])


def get_cond_images(args, zoomed_image, mag_level, lazy=False):
    """sample_ultra_res.py:304-400.  Returns (cond_images [P, 3|6, 1024, 1024], patch_pos, num_patches_width); with
    lazy=True cond_images is a `CondBank` (same indexing, windows computed on demand)."""
    patch_width = get_patch_width(args, mag_level)
    patch_dist = int(patch_width * (1 - args.overlap))
    W = zoomed_image.shape[3]
    num_patches_width = 1 + math.ceil((W - patch_width) / patch_dist)
    airs = getattr(args, "version", None) == "airs"
    if airs:
        num_patches_width = max(1, num_patches_width - 1)
    if mag_level == 2:
        patch_pos = tissue_patches(args, zoomed_image, patch_width, patch_dist, num_patches_width)
    else:
        patch_pos = [(i, j) for i in range(num_patches_width) for j in range(num_patches_width)]
    shifts = []
    for i, j in patch_pos:
        y, x = i * patch_dist, j * patch_dist
        shifts.append((W // 2 - (y + patch_width // 2), W // 2 - (x + patch_width // 2)))
    bank = CondBank(zoomed_image, shifts, 0.0 if airs else 0.95, 6 if getattr(args, "version", None) == "v2" else 3, patch_width)
    return (bank if lazy else bank.materialize()), patch_pos, num_patches_width


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


# ------------------------------------------------------------------------------------------------ round-synchronous schedule (NCCL fallback)
class Schedule:
    def __init__(self, rounds, owner):
        self.rounds, self.owner = rounds, owner  # rounds: list of {rank: [patch index, ...]}


def build_schedule(patch_pos, orientation, world, max_batch=1):
    """Single-stage, round-synchronous list schedule (round 1's scheduler; still used when the CUDA-IPC mailbox is not
    available and strips travel as grouped NCCL send/recv between rounds).  A patch is ready once its (up to three)
    neighbours that exist in the grid are done (sample_ultra_res.py:99-107).  Each round takes up to world * max_batch
    ready patches, longest-remaining-dependency-chain first, spread evenly; a patch prefers the rank that produced its
    `above` neighbour."""
    patch_pos = [tuple(p) for p in patch_pos]
    index = {p: k for k, p in enumerate(patch_pos)}
    deps = [[index[nb] for nb in neighbours(p, orientation).values() if nb in index] for p in patch_pos]
    users = [[index[q] for q in dependents(p, orientation).values() if q in index] for p in patch_pos]
    order = sorted(range(len(patch_pos)), key=lambda k: (-patch_pos[k][0], orientation * patch_pos[k][1]))
    prio = [0] * len(patch_pos)
    for k in order:
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


def plan_from_schedule(sched, stage, world):
    """Round-synchronous schedule expressed as a grid_plan.Plan (round r = simulated time r)."""
    batches, owner = [], {}
    for t, rnd in enumerate(sched.rounds):
        for r in sorted(rnd):
            batches.append(grid_plan.Batch(stage, list(rnd[r]), r, float(t), float(t + 1)))
            for k in rnd[r]:
                owner[(stage, k)] = r
    return grid_plan.Plan(batches, owner, float(len(sched.rounds)), (stage,), world, "rounds")


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
    {key: tensor} of what this rank received (NCCL fallback transport and the final gather)."""
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


_RUN_SEED = None


def run_seed(args):
    """Noise seed of this process group's run: args.seed when given, else one random draw (torch's global generator, so
    torch.manual_seed makes runs reproducible) on rank 0, broadcast -- every rank must key its patches' noise identically.
    The reference draws from unseeded per-process generators; this is what makes two runs produce different images."""
    global _RUN_SEED
    seed = getattr(args, "seed", None)
    if seed is not None:
        return int(seed)
    if _RUN_SEED is None:
        dist, rank, world = _dist()
        s = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)
        if world > 1:
            dev = _device_for(args, rank)
            s = s.to(dev)
            dist.broadcast(s, 0)
        _RUN_SEED = int(s.item())
    return _RUN_SEED


# ------------------------------------------------------------------------------------------------ hooks (tests / bench)
def default_model_provider(mag_level, unet_number, device, args):
    """load_model of the reference (sample_ultra_res.py:36-65; outpainting.py's variant for mag_level == 'outpaint'):
    factory by --version, checkpoint from args, strict load with a restore_parts fallback.  Models are cached per
    (mag, stage): they stay resident instead of being reloaded."""
    from .trainer import __version__, restore_parts

    if mag_level == "outpaint":  # outpainting.py:24-45: the unconditional cascade, checkpoints args.unet{n}
        from .factories import init_imagen_uncond

        imagen = init_imagen_uncond(unet_number, device=device)
        path = vars(args)[f"unet{unet_number}"]
    else:
        from .factories import init_imagen_ultra_res

        imagen = init_imagen_ultra_res(mag_level, unet_number, device=device, version=getattr(args, "version", None) or "")
        path = vars(args)[f"unet{unet_number}_mag{mag_level}"]
    try:
        from fsspec.core import url_to_fs

        fs, _ = url_to_fs(path)
        with fs.open(path) as f:
            loaded = torch.load(f, map_location="cpu")
    except ImportError:
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


def load_model(mag_level, unet_number, device, args, provider=None):
    provider = provider or MODEL_PROVIDER
    key = (mag_level, unet_number, str(device), id(provider))
    if key not in _MODEL_CACHE:
        _MODEL_CACHE[key] = provider(mag_level, unet_number, device, args)
    model = _MODEL_CACHE[key]
    precision = getattr(args, "precision", None)  # "fp32": the precise CUDA-core path of every UNet (validation runs)
    if precision is not None and hasattr(model, "set_precision"):
        model.set_precision(precision)
    return model


def _device_for(args, rank):
    dev = getattr(args, "device", None)
    if dev is not None:
        return torch.device(dev)
    return torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))


# ------------------------------------------------------------------------------------------------ a9 / a10: plan execution
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


def _messages(plan, patch_pos, orientation, overlap, prev_owner, prev_stage):
    """Every cross-rank message of a plan, grouped by receiving rank: {dst: [((stage, k, kind), shape), ...]} in a
    deterministic order (identical on all ranks).  kinds: above / side / corner strips of a finished patch-stage, 'lowres'
    = a finished patch of stage u needed by stage u+1 of the same patch on another rank."""
    index = {p: k for k, p in enumerate(patch_pos)}
    incoming = {}
    for u in plan.stages:
        S = PATCH_SIZES[u]
        ov = int(overlap * S)
        for k, p in enumerate(patch_pos):
            src = plan.owner[(u, k)]
            for kind, q in dependents(p, orientation).items():
                if q in index and plan.owner[(u, index[q])] != src:
                    incoming.setdefault(plan.owner[(u, index[q])], []).append(((u, k, kind), _strip_shape(kind, S, ov)))
            if (u + 1) in plan.stages and plan.owner[(u + 1, k)] != src:
                incoming.setdefault(plan.owner[(u + 1, k)], []).append(((u, k, "lowres"), (3, S, S)))
    if prev_owner is not None:  # outputs of an earlier single-stage call that live on another rank than their consumer
        u, S = plan.stages[0], PATCH_SIZES[prev_stage]
        for k in range(len(patch_pos)):
            if prev_owner[k] != plan.owner[(u, k)]:
                incoming.setdefault(plan.owner[(u, k)], []).append(((prev_stage, k, "lowres"), (3, S, S)))
    return incoming


def _make_plan(patch_pos, orientation, world, stages, args, imagens):
    """args.cost_table ({stage: {B: ms per step}}, identical on every rank) overrides the built-in cost model; args.plan_policy
    pins the packing variant; args.max_batch caps the batch (int or {stage: cap})."""
    steps = {}
    for u in stages:
        try:
            steps[u] = int(imagens[u].noise_schedulers[u - 1].num_timesteps)
        except Exception:  # noqa: BLE001 -- stand-in models of the CPU tests
            steps[u] = grid_plan.FULL_STEPS[u]
    mb = getattr(args, "max_batch", None) or MAX_BATCH
    n = len(patch_pos)
    mb = {u: max(1, min(mb if isinstance(mb, int) else mb.get(u, MAX_BATCH[u]), n)) for u in (1, 2, 3)}
    return grid_plan.build_plan(patch_pos, orientation, world, stages=stages, steps=steps, resample=max(1, int(args.inpaint_resample)),
                                max_batch=mb, policy=getattr(args, "plan_policy", None), table=getattr(args, "cost_table", None))


def _run(mag_level, stages, args, lowres_image, cond_image, patch_pos, overlap, orientation, num_patches_width, provider=None):
    """Executes `stages` (consecutive unet numbers) over the grid and returns the PatchSet of the last one."""
    dist, rank, world = _dist()
    device = _device_for(args, rank)
    t_wall = time.time()
    imagens = {u: load_model(mag_level, u, device, args, provider) for u in stages}
    seed = run_seed(args)
    for im in imagens.values():
        if hasattr(im, "noise_seed"):
            im.noise_seed = seed
    n = len(patch_pos) if patch_pos is not None else (cond_image.shape[0] if cond_image is not None else 1)
    sample_kw = lambda u: dict(return_pil_images=False, start_at_unet_number=u, stop_at_unet_number=u,
                               inpaint_resample_times=args.inpaint_resample, use_tqdm=False, device=device)
    if isinstance(cond_image, CondBank) and device.type == "cuda":
        cond_image = cond_image.to(device)

    if patch_pos is None:  # magnification 0: plain sampling, replicated on every rank (deterministic noise)
        prev = lowres_image
        for u in stages:
            out = []
            for idx in range(n):
                lr = None if prev is None else prev[idx].to(device)
                ci = None if cond_image is None else cond_image[idx].unsqueeze(0)
                out.append(imagens[u].sample(batch_size=1, cond_images=ci, start_image_or_video=lr, inpaint_images=None, inpaint_masks=None,
                                             noise_key=[_noise_key(mag_level, idx)], **sample_kw(u)))
            prev = out
        return PatchSet(prev, owner=None)

    patch_pos = [tuple(p) for p in patch_pos]
    index = {p: k for k, p in enumerate(patch_pos)}
    prev_owner = getattr(lowres_image, "owner", None) if (lowres_image is not None and world > 1) else None
    prev_stage = stages[0] - 1

    # ---- plan + transport (collective decisions: every rank takes the same branch)
    transport, plan = None, None
    if world > 1 and device.type == "cuda":
        plan = _make_plan(patch_pos, orientation, world, stages, args, imagens)
        from .grid_exec import cached_peer_mailbox

        transport = cached_peer_mailbox(dist, rank, world, device, _messages(plan, patch_pos, orientation, overlap, prev_owner, prev_stage))
        if transport is None:  # NCCL fallback: stage-major, round-synchronous
            prev = lowres_image
            for u in stages:
                prev = _run_stage_rounds(mag_level, u, args, prev, cond_image, patch_pos, overlap, orientation, num_patches_width, imagens[u])
            return prev
    elif world > 1:
        from .grid_exec import TaggedTransport

        plan = _make_plan(patch_pos, orientation, world, stages, args, imagens)
        transport = TaggedTransport(dist, rank, n)
    else:
        plan = _make_plan(patch_pos, orientation, 1, stages, args, imagens)

    patch_width = get_patch_width(args, mag_level) if cond_image is not None else 0  # only used for fallback neighbour crops
    patch_dist = int(patch_width * (1 - overlap))
    owner = plan.owner
    results = {u: {} for u in stages}
    if prev_owner is not None:  # hand earlier-stage outputs to the ranks that consume them
        Sp = PATCH_SIZES[prev_stage]
        for k in range(n):
            dst = owner[(stages[0], k)]
            if prev_owner[k] == rank and dst != rank:
                transport.post(dst, (prev_stage, k, "lowres"), lowres_image[k].to(device).reshape(3, Sp, Sp))

    trace = [] if (os.environ.get("KD_GRID_TRACE") and device.type == "cuda") else None  # per-batch CUDA events: inputs / sample / post
    for batch in plan.for_rank(rank):
        u, mine = batch.stage, batch.patches
        S = PATCH_SIZES[u]
        ov = int(overlap * S)
        inpaints, masks = [], []
        if trace is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            t_host0 = time.time()
        for k in mine:
            i, j = patch_pos[k]
            strips = {}
            for kind, nb in neighbours((i, j), orientation).items():
                if nb in index:
                    kk = index[nb]
                    if owner[(u, kk)] == rank:
                        strips[kind] = _strip_of(results[u][kk], kind, S, ov, orientation)
                    else:
                        g = transport.fetch(owner[(u, kk)], (u, kk, kind), _strip_shape(kind, S, ov), device)
                        strips[kind] = (g, g.shape[1] * g.shape[2], g.shape[2])
                else:
                    fb = _fallback_neighbour(kind, (i, j), orientation, num_patches_width, None if cond_image is None else cond_image[k],
                                             patch_width, patch_dist, S, device)
                    strips[kind] = None if fb is None else _strip_of(fb, kind, S, ov, orientation)
            ip, im = CANVAS_FN(S, ov, orientation, strips["above"], strips["side"], strips["corner"], device)
            inpaints.append(ip)
            masks.append(im)
        lr = None
        if u > stages[0]:  # previous stage of the same patches, produced inside this plan
            Sp = PATCH_SIZES[u - 1]
            lr = torch.cat([results[u - 1][k] if owner[(u - 1, k)] == rank
                            else transport.fetch(owner[(u - 1, k)], (u - 1, k, "lowres"), (3, Sp, Sp), device).reshape(1, 3, Sp, Sp)
                            for k in mine], 0)
        elif lowres_image is not None:
            Sp = PATCH_SIZES[prev_stage]
            lr = torch.cat([lowres_image[k].to(device, non_blocking=True) if (prev_owner is None or prev_owner[k] == rank)
                            else transport.fetch(prev_owner[k], (prev_stage, k, "lowres"), (3, Sp, Sp), device).reshape(1, 3, Sp, Sp)
                            for k in mine], 0)
        ci = None if cond_image is None else torch.stack([cond_image[k] for k in mine]).to(device)
        if trace is not None:
            ev[1].record()
            t_host1 = time.time()
        out = imagens[u].sample(batch_size=len(mine), cond_images=ci, start_image_or_video=lr, inpaint_images=torch.stack(inpaints),
                                inpaint_masks=torch.stack(masks), noise_key=[_noise_key(mag_level, k) for k in mine], **sample_kw(u))
        if trace is not None:
            ev[2].record()
            t_host2 = time.time()
        for b, k in enumerate(mine):
            results[u][k] = out[b:b + 1].contiguous()
        if transport is not None:  # push what dependents on other ranks need, as soon as it exists
            for k in mine:
                for kind, q in dependents(patch_pos[k], orientation).items():
                    if q in index and owner[(u, index[q])] != rank:
                        transport.post(owner[(u, index[q])], (u, k, kind), _strip_of(results[u][k], kind, S, ov, orientation)[0])
                if (u + 1) in plan.stages and owner[(u + 1, k)] != rank:
                    transport.post(owner[(u + 1, k)], (u, k, "lowres"), results[u][k].reshape(3, S, S))
        if trace is not None:
            ev[3].record()
            trace.append((u, len(mine), ev, (t_host1 - t_host0, t_host2 - t_host1, time.time() - t_host2)))
    if transport is not None:
        transport.finish()
    last = stages[-1]
    out = PatchSet([results[last].get(k) for k in range(n)], owner={k: owner[(last, k)] for k in range(n)})
    LAST_RUN.clear()
    LAST_RUN.update(stages=list(stages), world=world, plan_makespan_s=plan.makespan, plan_policy=str(plan.policy), batches=len(plan.batches),
                    batch_sizes={str(u): v for u, v in plan.batch_sizes().items()}, plan_busy_fraction=plan.busy_fraction(),
                    transport="none (single rank)" if transport is None else transport.name,
                    bytes_sent_this_rank=0 if transport is None else transport.bytes_sent, host_wall_s=time.time() - t_wall)
    if trace is not None:  # device time of the three phases (waiting for neighbours' strips counts as "inputs") and host time spent issuing them
        torch.cuda.synchronize(device)
        tot = dict(batches=len(trace), inputs_ms=0.0, sample_ms=0.0, post_ms=0.0, host_inputs_ms=0.0, host_sample_ms=0.0, host_post_ms=0.0)
        for _, _, ev, host in trace:
            tot["inputs_ms"] += ev[0].elapsed_time(ev[1])
            tot["sample_ms"] += ev[1].elapsed_time(ev[2])
            tot["post_ms"] += ev[2].elapsed_time(ev[3])
            tot["host_inputs_ms"] += host[0] * 1e3
            tot["host_sample_ms"] += host[1] * 1e3
            tot["host_post_ms"] += host[2] * 1e3
        tot["span_ms"] = trace[0][2][0].elapsed_time(trace[-1][2][3]) if trace else 0.0
        LAST_RUN["trace"] = tot
    out.plan = plan
    if transport is not None:
        if hasattr(transport, "recycle"):
            transport.recycle()  # cached mailbox: barrier + next epoch
        else:
            transport.close()
    return out


def _run_stage_rounds(mag_level, unet_number, args, lowres_image, cond_image, patch_pos, overlap, orientation, num_patches_width, imagen):
    """NCCL fallback (round 1's executor): one stage, round-synchronous; after each round the strips dependents on other
    ranks need travel as ONE grouped batch_isend_irecv."""
    dist, rank, world = _dist()
    device = _device_for(args, rank)
    S = PATCH_SIZES[unet_number]
    n = len(patch_pos)
    index = {p: k for k, p in enumerate(patch_pos)}
    mb = getattr(args, "max_batch", None) or {1: 16, 2: 8, 3: 2}
    mb = mb if isinstance(mb, int) else mb.get(unet_number, 2)
    sched = build_schedule(patch_pos, orientation, world, max(1, min(mb, n)))
    owner = sched.owner
    overlap_pos = int(overlap * S)
    patch_width = get_patch_width(args, mag_level) if cond_image is not None else 0
    patch_dist = int(patch_width * (1 - overlap))
    sample_kw = dict(return_pil_images=False, start_at_unet_number=unet_number, stop_at_unet_number=unet_number,
                     inpaint_resample_times=args.inpaint_resample, use_tqdm=False, device=device)
    lowres_local = {}
    if lowres_image is not None:
        prev_owner = getattr(lowres_image, "owner", None)
        if prev_owner is None:
            lowres_local = {k: lowres_image[k] for k in range(n) if owner[k] == rank}
        else:
            s_prev = PATCH_SIZES[unet_number - 1]
            moves = [(prev_owner[k], owner[k], k, None if lowres_image[k] is None else lowres_image[k].to(device), (1, 3, s_prev, s_prev))
                     for k in range(n)]
            lowres_local = _exchange(moves, rank, device)
    results, ghosts, sent = {}, {}, 0
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
        transfers = []
        for r in sorted(rnd):
            for k in rnd[r]:
                for kind, q in dependents(patch_pos[k], orientation).items():
                    if q in index and owner[index[q]] != r:
                        t = _strip_of(results[k], kind, S, overlap_pos, orientation)[0] if r == rank else None
                        sent += 0 if t is None else t.numel() * 4
                        transfers.append((r, owner[index[q]], (k, kind), t, _strip_shape(kind, S, overlap_pos)))
        ghosts.update(_exchange(transfers, rank, device))
    LAST_RUN.clear()
    LAST_RUN.update(stages=[unet_number], world=world, rounds=len(sched.rounds), transport="NCCL grouped send/recv between rounds (fallback)",
                    bytes_sent_this_rank=sent)
    out = PatchSet([results.get(k) for k in range(n)], owner=dict(owner))
    out.plan = plan_from_schedule(sched, unet_number, world)
    return out


def generate_image_with_unet(mag_level, unet_number, args, lowres_image, cond_image, patch_pos, overlap, orientation, num_patches_width,
                             provider=None):
    """sample_ultra_res.py:213-261 (dispatcher) + :75-210 (worker), SPMD: one stage over the grid.  Returns a PatchSet of
    (1,3,S,S) tensors (entry k is None on ranks that do not own patch k; `.owner[k]` names the rank that does)."""
    return _run(mag_level, (unet_number,), args, lowres_image, cond_image, patch_pos, overlap, orientation, num_patches_width, provider)


# ------------------------------------------------------------------------------------------------ a11 / a15
def generate_image(mag_level, args, cond_image=None, patch_pos=None, overlap=0.25, orientation=-1, num_patches_width=1, lowres_image=None,
                   provider=None):
    """sample_ultra_res.py:264-270: cascade 64 -> 256 -> 1024 over all patches.  The reference is stage-major; here the three
    stages form ONE dependency-driven plan (same results, see grid_plan.py).  args.stage_major=True restores the reference's
    order (three separate single-stage plans)."""
    stages = (1, 2, 3) if lowres_image is None else (2, 3)
    if getattr(args, "stage_major", False):
        prev = lowres_image
        for u in stages:
            prev = generate_image_with_unet(mag_level, u, args, prev, cond_image, patch_pos, overlap, orientation, num_patches_width, provider)
        return prev
    return _run(mag_level, stages, args, lowres_image, cond_image, patch_pos, overlap, orientation, num_patches_width, provider)


def gather_patches(patches, dst=0):
    """Bring every patch of a distributed PatchSet to rank `dst` (as CPU tensors, like the reference's results)."""
    dist, rank, world = _dist()
    owner = getattr(patches, "owner", None)
    if world == 1 or owner is None:
        return [p.cpu() for p in patches]
    local = next((p for p in patches if p is not None), None)
    if local is not None:
        device = local.device
    else:  # a rank that owns no patch (world > number of patches) still takes part in the exchange
        device = torch.device("cpu") if dist.get_backend() == "gloo" else torch.device("cuda", torch.cuda.current_device())
    # every rank must post receives of the right shape: the owner of patch 0 announces it
    meta = torch.tensor(list(patches[0].shape) if owner[0] == rank else [0, 0, 0, 0], dtype=torch.int64, device=device)
    dist.broadcast(meta, owner[0])
    shape = tuple(int(v) for v in meta.tolist())
    moves = [(owner[k], dst, k, patches[k], shape) for k in range(len(patches))]
    got = _exchange(moves, rank, device)
    return [got[k].cpu() for k in range(len(patches))] if rank == dst else None


def _cell_index(patch_pos, n, device):
    cell = torch.full((n * n,), -1, dtype=torch.int32)
    for k, (i, j) in enumerate(patch_pos):
        cell[i * n + j] = k
    return cell.to(device)


def stitch_device(zoomed_image, patches, patch_pos, num_patches_width, overlap, device):
    """sample_ultra_res.py:430-446 on the GPU(s): returns the stitched (1,3,Wc,Wc) image on EVERY rank's device.
    Each rank pastes the pixels its patches own (see kd_patch_paste) into every rank's canvas -- its own directly, the
    others' through their CUDA-IPC mapping -- and each canvas' background comes from the local bilinear kernel."""
    from . import ops

    dist, rank, world = _dist()
    patch_pos = [tuple(p) for p in patch_pos]
    n = num_patches_width
    patch_dist = int(PATCH_SIZE * (1 - overlap))
    Wc = PATCH_SIZE + (n - 1) * patch_dist
    cell = _cell_index(patch_pos, n, device)
    zoomed = None if zoomed_image is None else zoomed_image[0].to(device).float().contiguous()
    owner = getattr(patches, "owner", None)
    if world == 1 or owner is None:
        canvas = torch.empty((1, 3, Wc, Wc), device=device, dtype=torch.float32)
        ops.canvas_fill(zoomed, canvas.data_ptr(), Wc, cell, n, patch_dist, PATCH_SIZE)
        for k, (i, j) in enumerate(patch_pos):
            ops.patch_paste(patches[k].to(device).reshape(3, PATCH_SIZE, PATCH_SIZE).contiguous(), canvas.data_ptr(), Wc, cell, n, patch_dist, k, i, j)
        return canvas
    from .grid_exec import _CANVAS_CACHE, cached_peer_mailbox

    incoming = {r: [(("canvas",), (3, Wc, Wc))] + [(("done", s), (1, 1, 1)) for s in range(world) if s != r] for r in range(world)}
    box = cached_peer_mailbox(dist, rank, world, device, incoming, max_entries=1, cache=_CANVAS_CACHE)
    if box is None:  # NCCL fallback: gather on rank 0, stitch there, broadcast
        full = gather_patches(patches)
        canvas = torch.empty((1, 3, Wc, Wc), device=device, dtype=torch.float32)
        if rank == 0:
            ops.canvas_fill(zoomed, canvas.data_ptr(), Wc, cell, n, patch_dist, PATCH_SIZE)
            for k, (i, j) in enumerate(patch_pos):
                ops.patch_paste(full[k].to(device).reshape(3, PATCH_SIZE, PATCH_SIZE).contiguous(), canvas.data_ptr(), Wc, cell, n, patch_dist, k, i, j)
        dist.broadcast(canvas, 0)
        return canvas
    off = {r: box.layout[r][0][("canvas",)][1] for r in range(world)}
    ops.canvas_fill(zoomed, box.base[rank] + off[rank], Wc, cell, n, patch_dist, PATCH_SIZE)
    token = torch.ones(1, 1, 1, device=device, dtype=torch.float32)
    for k, (i, j) in enumerate(patch_pos):
        if owner[k] == rank:
            p = patches[k].reshape(3, PATCH_SIZE, PATCH_SIZE)
            for r in range(world):
                ops.patch_paste(p, box.base[r] + off[r], Wc, cell, n, patch_dist, k, i, j)
    for r in range(world):
        if r != rank:
            box.post(r, ("done", rank), token)
    for s in range(world):
        if s != rank:
            box.fetch(s, ("done", s), (1, 1, 1), device)
    own = box.own[off[rank] // 4: off[rank] // 4 + 3 * Wc * Wc].view(1, 3, Wc, Wc)
    canvas = own.clone()
    box.finish()
    box.recycle()  # barrier: every rank has cloned its canvas before anybody pastes the next image into it
    return canvas


def generate_high_res_image(zoomed_image, mag_level, args, provider=None):
    """sample_ultra_res.py:415-448.  Returns the stitched image on every rank (on the GPU when sampling ran there, so the
    next magnification level -- `generate_high_res_image(mag1_full_image, 2, args)` in the reference's main() -- chains without
    leaving the device)."""
    dist, rank, world = _dist()
    device = _device_for(args, rank)
    t0 = time.time()
    zoomed_dev = zoomed_image.to(device) if device.type == "cuda" else zoomed_image
    cond_images, patch_pos, num_patches_width = get_cond_images(args, zoomed_dev, mag_level, lazy=True)
    patch_width = get_patch_width(args, mag_level)
    if getattr(args, "ignore_unet_1", False):
        top = int(round((PATCH_SIZE - patch_width) / 2.0))
        lowres_image = [c[:, top:top + patch_width, top:top + patch_width].unsqueeze(0) for c in cond_images]
    else:
        lowres_image = None
    orientation = choose_orientation(patch_pos)
    t1 = time.time()
    mag_images = generate_image(mag_level, args, cond_image=cond_images, patch_pos=patch_pos, overlap=args.overlap, orientation=orientation,
                                num_patches_width=num_patches_width, lowres_image=lowres_image, provider=provider)
    if device.type == "cuda":
        torch.cuda.synchronize(device)
    t2 = time.time()
    if device.type == "cuda":
        full = stitch_device(zoomed_dev, mag_images, patch_pos, num_patches_width, args.overlap, device)
        torch.cuda.synchronize(device)
    else:  # host path of the CPU tests (stand-in models): gather + the reference's paste loop, then share the result
        gathered = gather_patches(mag_images)
        full = stitch(zoomed_image, gathered, patch_pos, num_patches_width, args.overlap) if gathered is not None else None
        if world > 1:
            box = [full]
            dist.broadcast_object_list(box, 0)
            full = box[0]
    LAST_RUN.update(cond_images_s=t1 - t0, sampling_s=t2 - t1, stitch_s=time.time() - t2, patches=len(patch_pos))
    return full


def stitch(zoomed_image, mag_images, patch_pos, num_patches_width, overlap):
    """sample_ultra_res.py:430-446 exactly as the reference runs it (host): canvas = bilinear upsample of the zoomed image;
    patches pasted row-major, later patches overwrite earlier ones, no blending.  The GPU path is `stitch_device`."""
    patch_dist = int(PATCH_SIZE * (1 - overlap))
    full_image_width = PATCH_SIZE + (num_patches_width - 1) * patch_dist
    full_image = F.interpolate(zoomed_image.cpu().float(), size=(full_image_width, full_image_width), mode="bilinear", align_corners=False)
    for index, (i, j) in enumerate(patch_pos):
        y, x = i * patch_dist, j * patch_dist
        full_image[0, :, y:y + PATCH_SIZE, x:x + PATCH_SIZE] = mag_images[index][0].cpu()
    return full_image

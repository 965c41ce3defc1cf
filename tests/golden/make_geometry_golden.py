"""Generates tests/golden/geometry_golden.json by importing the REFERENCE's own patch-grid code
(/root/reference/sample_ultra_res.py) on the CPU with its missing third-party modules stubbed in sys.modules.

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_geometry_golden.py

Nothing of the reference is copied: the fixture holds only OUTPUTS of its functions on seeded synthetic inputs
(patch coordinates, grid sizes, orientation, SHA-256 digests of the tensors it builds, the per-patch inpaint masks as
run-length rows, and the processing order of its worker loop).
"""
import hashlib
import json
import math
import os
import queue
import sys
import types

import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "geometry_golden.json")


def stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("imagen_pytorch", Unet=object, ImagenTrainer=object, Imagen=object, NullUnet=object, SRUnet1024=object, ElucidatedImagen=object)
    mod("imagen_pytorch.trainer", restore_parts=lambda a, b: a)
    mod("imagen_pytorch.version", __version__="1.18.5")
    sk = mod("skimage")
    sk.color = mod("skimage.color", rgb2hsv=None)
    mod("skimage.io")
    mod("skimage.transform")
    fs = mod("fsspec")
    fs.core = mod("fsspec.core", url_to_fs=None)
    mod("h5py")
    mod("slideio")
    mp = mod("matplotlib")
    mp.pyplot = mod("matplotlib.pyplot")
    mp.cm = mod("matplotlib.cm")
    mod("joblib", Parallel=None, delayed=None)


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()


class Args:
    def __init__(self, **kw):
        self.version, self.overlap, self.inpaint_resample, self.ignore_unet_1, self.num_gpus = "v_param", 0.25, 1, False, 1
        self.__dict__.update(kw)


class CpuTorch:
    """Proxy of the torch module whose device() always answers CPU (the worker does .to(torch.device('cuda:rank')))."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*a, **k):
        return torch.device("cpu")


class FakeImagen:
    """Stands in for the model: returns a deterministic image per call and records the inpainting inputs."""

    def __init__(self, S, log):
        self.S, self.log, self.calls = S, log, 0

    def sample(self, batch_size, return_pil_images, cond_images, start_image_or_video, start_at_unet_number, stop_at_unet_number,
               inpaint_images, inpaint_masks, inpaint_resample_times, use_tqdm, device):
        # keyed by the patch's conditioning image, not by call order: any valid processing order gives the same patches
        key = int(cond_images.double().sum().item() * 1e3) % (2 ** 31)
        g = torch.Generator().manual_seed(key)
        out = torch.rand(1, 3, self.S, self.S, generator=g)
        if inpaint_images is not None:  # like the real sampler, known pixels are pasted back
            m = inpaint_masks.bool()[:, None]
            out = out * ~m + inpaint_images * m
        self.log.append(dict(call=self.calls, key=key, inpaint=None if inpaint_images is None else sha(inpaint_images),
                             mask_rows=None if inpaint_masks is None else mask_rle(inpaint_masks[0]),
                             mask_sum=None if inpaint_masks is None else int(inpaint_masks.sum().item())))
        self.calls += 1
        return out


def mask_rle(mask):
    """Per-row (first, count) of ones -- the reference's masks are unions of a top band and a side band."""
    rows = []
    for r in mask:
        nz = torch.nonzero(r).flatten()
        rows.append([int(nz[0]), int(nz.numel())] if nz.numel() else [0, 0])
    # compress identical consecutive rows
    out, prev, n = [], None, 0
    for r in rows:
        if r == prev:
            n += 1
        else:
            if prev is not None:
                out.append(prev + [n])
            prev, n = r, 1
    out.append(prev + [n])
    return out


def synthetic_slide(W, seed, dark_background=False):
    """Slide-like test image: near-neutral background, a few purple tissue blobs, and 3x3 specks the 5x5 erosion must remove."""
    g = torch.Generator().manual_seed(seed)
    bg = 0.02 if dark_background else 0.95
    img = torch.full((1, 3, W, W), bg) + (torch.rand(1, 1, W, W, generator=g) - 0.5) * 0.004
    yy, xx = torch.meshgrid(torch.arange(W), torch.arange(W), indexing="ij")
    colour = torch.tensor([0.80, 0.50, 0.80]).view(1, 3, 1, 1)
    for _ in range(4):
        cy, cx = (int(v) for v in torch.randint(100, W - 100, (2,), generator=g))
        ry, rx = (int(v) for v in torch.randint(30, 140, (2,), generator=g))
        m = (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) <= 1.0
        img = torch.where(m[None, None], colour + (torch.rand(1, 3, W, W, generator=g) - 0.5) * 0.05, img)
    for _ in range(30):
        cy, cx = (int(v) for v in torch.randint(2, W - 3, (2,), generator=g))
        img[:, :, cy - 1:cy + 2, cx - 1:cx + 2] = colour
    return img.clamp(0, 1).contiguous()


def main():
    import contextlib
    import io

    stub_modules()
    sys.path.insert(0, REF)
    import sample_ultra_res as R

    gold = {"source": "sample_ultra_res.py functions run on CPU with stubbed third-party imports", "cases": {}}
    C = gold["cases"]

    # a12 get_patch_width
    C["patch_width"] = {f"{v}/{m}": R.get_patch_width(Args(version=v), m) for v in ("v_param", "", "v2", "airs") for m in (1, 2)}

    # a13 get_cond_images (mag 1: no tissue filter)
    cond_cases = []
    for version, W, overlap, seed in [("v_param", 1024, 0.25, 1), ("v_param", 1024, 0.5, 2), ("v2", 1024, 0.25, 3), ("airs", 1024, 0.25, 4),
                                      ("", 700, 0.25, 5), ("v_param", 1190, 0.25, 6), ("v_param", 249, 0.5, 7),
                                      ("v_param", 1158, 0.25, 8)]:  # W=1158: row / column 4 has shift == 0 (the fill-everything quirk of :380-388)
        g = torch.Generator().manual_seed(seed)
        zoomed = torch.rand(1, 3, W, W, generator=g)
        args = Args(version=version, overlap=overlap)
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            cond, pos, n = R.get_cond_images(args, zoomed.clone(), 1)
        cond_cases.append(dict(version=version, W=W, overlap=overlap, seed=seed, shape=list(cond.shape), n=n, patch_pos=[list(p) for p in pos],
                               sha=sha(cond), first_sha=sha(cond[0]), last_sha=sha(cond[-1])))
    C["cond_images"] = cond_cases

    # a13 at magnification 2: the tissue filter (:317-352).  skimage is absent here, so the reference code runs with OpenCV's
    # RGB->HSV (an independent implementation; H scaled from degrees to [0,1]) standing in for skimage.color.rgb2hsv.
    import cv2
    import numpy as np

    R.color.rgb2hsv = lambda a: cv2.cvtColor(np.ascontiguousarray(a, dtype=np.float32), cv2.COLOR_RGB2HSV) / np.array([360.0, 1.0, 1.0], dtype=np.float32)
    tissue = []
    for version, W, seed in [("v_param", 1400, 21), ("airs", 1300, 22)]:
        zoomed = synthetic_slide(W, seed, dark_background=(version == "airs"))
        args = Args(version=version, overlap=0.25)
        with contextlib.redirect_stdout(io.StringIO()):
            cond, pos, n = R.get_cond_images(args, zoomed.clone(), 2)
        tissue.append(dict(version=version, W=W, seed=seed, n=n, patch_pos=[list(p) for p in pos], shape=list(cond.shape), sha=sha(cond),
                           first_sha=sha(cond[0]), last_sha=sha(cond[-1])))
    C["tissue"] = tissue

    # a14 get_next_patches + a15 orientation rule
    nxt = []
    full = [(i, j) for i in range(5) for j in range(5)]
    ragged = [(0, 2), (0, 3), (1, 1), (1, 2), (1, 3), (2, 0), (2, 1), (2, 2), (3, 0), (3, 1), (4, 4)]
    for name, patches in [("full5", full), ("ragged", ragged), ("single", [(0, 0)]), ("row", [(0, j) for j in range(4)])]:
        for o in (-1, 1):
            p, w = R.get_next_patches(patches, o)
            nxt.append(dict(name=name, patches=[list(x) for x in patches], orientation=o, roots=[list(x) for x in p], waiting=[list(x) for x in w]))
    C["next_patches"] = nxt

    # a9 worker loop: inpaint canvas + mask + processing order, driven with a fake model on the CPU
    worker = []
    for version, grid_pos, n_w, o, unet_number, overlap in [
        ("v_param", full[:9] and [(i, j) for i in range(3) for j in range(3)], 3, -1, 1, 0.25),
        ("v_param", [(i, j) for i in range(3) for j in range(3)], 3, 1, 2, 0.25),
        ("v_param", ragged, 5, -1, 1, 0.25),
        ("v_param", ragged, 5, 1, 1, 0.5),
    ]:
        S = R.PATCH_SIZES[unet_number]
        args = Args(version=version, overlap=overlap, inpaint_resample=2)
        log = []
        fake = FakeImagen(S, log)
        R.load_model = lambda *a, **k: fake
        R.torch = CpuTorch()
        g = torch.Generator().manual_seed(77)
        cond = torch.rand(len(grid_pos), 3, 1024, 1024, generator=g)
        outq, done = queue.Queue(), {}

        class WorkQueue(queue.Queue):  # poison pill only once every patch is done (re-queued items come after it otherwise)
            def get(self_q, *a, **k):
                return None if len(done) == len(grid_pos) else queue.Queue.get(self_q, *a, **k)

        inq = WorkQueue()
        for idx, pos in enumerate(grid_pos):
            inq.put((idx, None, cond[idx], pos))
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            R.generate_image_distributed(0, 1, unet_number, args, inq, outq, done, overlap, o, grid_pos, n_w)
        R.torch = torch
        order = []
        while not outq.empty():
            order.append(outq.get()[0])
        worker.append(dict(version=version, patch_pos=[list(p) for p in grid_pos], num_patches_width=n_w, orientation=o, unet_number=unet_number,
                           overlap=overlap, S=S, order=order, calls=log, outputs=[sha(done[i]) for i in range(len(grid_pos))], cond_seed=77))
    C["worker"] = worker

    # a15 stitch: generate_high_res_image with generate_image replaced by a fake that returns seeded patches
    stitch = []
    for W, overlap, seed in [(1024, 0.25, 11), (600, 0.5, 12)]:
        g = torch.Generator().manual_seed(seed)
        zoomed = torch.rand(1, 3, W, W, generator=g)
        args = Args(version="v_param", overlap=overlap)
        rec = {}

        def fake_generate_image(mag_level, args_, cond_image=None, patch_pos=None, overlap=0.25, orientation=-1, num_patches_width=1, lowres_image=None):
            rec.update(orientation=orientation, n=num_patches_width, patch_pos=[list(p) for p in patch_pos])
            out = []
            for k in range(len(patch_pos)):
                gg = torch.Generator().manual_seed(5000 + k)
                out.append(torch.rand(1, 3, 1024, 1024, generator=gg))
            return out

        R.generate_image = fake_generate_image
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            full_image = R.generate_high_res_image(zoomed.clone(), 1, args)
        stitch.append(dict(W=W, overlap=overlap, seed=seed, shape=list(full_image.shape), sha=sha(full_image), **rec))
    C["stitch"] = stitch

    with open(OUT, "w") as fh:
        json.dump(gold, fh, indent=1)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()

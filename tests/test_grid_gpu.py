"""Patch-grid sampler on the GPU: the CUDA canvas kernel inside the real scheduler, and invariance of every patch to how
patches are batched (the property that makes 1/2/4/8-GPU runs produce identical images)."""
import types

import pytest
import torch

from helpers import U1_KW, U2_KW

pytestmark = pytest.mark.gpu


def _provider(timesteps):
    from kidney_diffusion_b200.factories import FixedNullUnet, randomize_zero_init_
    from kidney_diffusion_b200 import Imagen, Unet

    def make(mag, n, device, args):
        torch.manual_seed(100 + n)
        kw1 = dict(U1_KW, cond_images_channels=3)
        unets = (Unet(**kw1) if n == 1 else FixedNullUnet(), Unet(**U2_KW) if n == 2 else FixedNullUnet(lowres_cond=True),
                 FixedNullUnet(lowres_cond=True))
        im = Imagen(unets=unets, image_sizes=(64, 256, 1024), timesteps=timesteps, pred_objectives=("noise", "v", "v"),
                    random_crop_sizes=(None, None, 256), condition_on_text=False)
        randomize_zero_init_(im)
        return im.to(device).eval()

    return make


def test_grid_stage_batch_invariance_and_borders(cuda_lib):
    from kidney_diffusion_b200 import grid

    grid._MODEL_CACHE.clear()
    grid.MODEL_PROVIDER = _provider((3, 2, 2))
    grid.CANVAS_FN = grid.default_canvas
    g = torch.Generator().manual_seed(0)
    zoomed = torch.rand(1, 3, 420, 420, generator=g)
    outs = {}
    for mb in (1, 4):
        args = types.SimpleNamespace(version="v_param", overlap=0.25, inpaint_resample=2, ignore_unet_1=False, num_gpus=1, device="cuda:0",
                                     max_batch=mb)
        cond, pos, n = grid.get_cond_images(args, zoomed, 1)
        assert n == 4 and len(pos) == 16
        o = grid.choose_orientation(pos)
        low = grid.generate_image_with_unet(1, 1, args, None, cond, pos, 0.25, o, n)
        med = grid.generate_image_with_unet(1, 2, args, low, cond, pos, 0.25, o, n)
        outs[mb] = (torch.cat(list(low)), torch.cat(list(med)))
    for a, b in zip(outs[1], outs[4]):
        assert torch.equal(a, b), "a patch must not depend on which patches share its batch"
    low, med = outs[1]
    S, ov = 256, 64
    # the overlap border of every patch equals its neighbour's facing strip (inpainted region is pasted back exactly)
    for k, (i, j) in enumerate(pos):
        if (i - 1, j) in pos:
            up = pos.index((i - 1, j))
            assert torch.equal(med[k, :, :ov, :][:, :, ov:], med[up, :, -ov:, :][:, :, ov:]) or torch.allclose(
                med[k, :, :ov, ov:], med[up, :, -ov:, ov:], atol=1e-6)
    assert float(med.min()) >= 0 and float(med.max()) <= 1


# ------------------------------------------------------------------------------------------------ N2 / N3 kernels vs the reference's golden outputs
import hashlib
import json
import os

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "geometry_golden.json")))["cases"]


def _sha(t):
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()


def _args(**kw):
    d = dict(version="v_param", overlap=0.25, inpaint_resample=1, ignore_unet_1=False, num_gpus=1, device="cuda:0", seed=3)
    d.update(kw)
    return types.SimpleNamespace(**d)


@pytest.mark.parametrize("case", GOLD["cond_images"], ids=lambda c: f"{c['version'] or 'default'}-W{c['W']}-ov{c['overlap']}")
def test_cond_gather_kernel_bit_exact_vs_reference(cuda_lib, case):
    """kd_cond_gather against outputs of the reference's own get_cond_images (roll + fill + CenterCrop [+ v2 nearest crop]),
    including W < 1024 (zero padding) and W = 1158 (shift == 0 fills the whole window)."""
    from kidney_diffusion_b200 import grid

    zoomed = torch.rand(1, 3, case["W"], case["W"], generator=torch.Generator().manual_seed(case["seed"]))
    cond, pos, n = grid.get_cond_images(_args(version=case["version"], overlap=case["overlap"]), zoomed.cuda(), 1)
    assert cond.is_cuda and n == case["n"] and [list(p) for p in pos] == case["patch_pos"] and list(cond.shape) == case["shape"]
    assert _sha(cond[0]) == case["first_sha"] and _sha(cond[-1]) == case["last_sha"] and _sha(cond) == case["sha"]


@pytest.mark.parametrize("case", GOLD["stitch"], ids=lambda c: f"W{c['W']}")
def test_stitch_kernels_bit_exact_vs_reference(cuda_lib, case):
    from kidney_diffusion_b200 import grid

    zoomed = torch.rand(1, 3, case["W"], case["W"], generator=torch.Generator().manual_seed(case["seed"]))
    patches = [torch.rand(1, 3, 1024, 1024, generator=torch.Generator().manual_seed(5000 + k)) for k in range(len(case["patch_pos"]))]
    full = grid.stitch_device(zoomed, grid.PatchSet([p.cuda() for p in patches]), [tuple(p) for p in case["patch_pos"]], case["n"],
                              case["overlap"], torch.device("cuda:0"))
    assert list(full.shape) == case["shape"] and _sha(full) == case["sha"]


def test_stitch_kernels_ragged_grid_background(cuda_lib):
    """Cells without a patch show the bilinear background (fp32 rounding of another implementation: 1e-6), covered pixels
    are bit-exact, later patches win the overlaps."""
    from kidney_diffusion_b200 import grid

    pos = [(0, 1), (1, 0), (1, 1), (2, 2)]
    zoomed = torch.rand(1, 3, 300, 300, generator=torch.Generator().manual_seed(1))
    patches = [torch.rand(1, 3, 1024, 1024, generator=torch.Generator().manual_seed(k)) for k in range(len(pos))]
    want = grid.stitch(zoomed, patches, pos, 3, 0.25)
    got = grid.stitch_device(zoomed, grid.PatchSet([p.cuda() for p in patches]), pos, 3, 0.25, torch.device("cuda:0")).cpu()
    covered = torch.zeros(want.shape[-2:], dtype=torch.bool)
    for i, j in pos:
        covered[i * 768:i * 768 + 1024, j * 768:j * 768 + 1024] = True
    assert torch.equal(got[0][:, covered], want[0][:, covered])
    assert torch.allclose(got, want, atol=2e-6, rtol=0)
    zero = grid.stitch_device(None, grid.PatchSet([p.cuda() for p in patches]), pos, 3, 0.25, torch.device("cuda:0")).cpu()
    assert float(zero[0][:, ~covered].abs().max()) == 0.0 and torch.equal(zero[0][:, covered], want[0][:, covered])


def test_strip_push_and_flag_wait(cuda_lib):
    """The mailbox primitives inside one process: strided strip -> contiguous slot + flag; a wait on a raised flag returns, a
    wait on a flag nobody raises gives up after its timeout and reports it (the GPU never hangs)."""
    from kidney_diffusion_b200 import ops

    patch = torch.rand(3, 64, 64, device="cuda")
    view = patch[:, 48:, 10:30]
    slot = torch.zeros(3 * 16 * 20, device="cuda")
    flags = torch.zeros(4, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.strip_push(view, view.stride(0), view.stride(1), slot.data_ptr(), flags.data_ptr() + 4, 1)
    ops.flag_wait(flags.data_ptr() + 4, 1, 5.0, status)
    torch.cuda.synchronize()
    assert torch.equal(slot.view(3, 16, 20), view) and flags.tolist() == [0, 1, 0, 0] and int(status.item()) == 0
    ops.flag_wait(flags.data_ptr() + 8, 1, 0.05, status)
    torch.cuda.synchronize()
    assert int(status.item()) == 1


# ------------------------------------------------------------------------------------------------ peer mailbox across processes
def _two_rank_worker(rank, world, port, ret):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)  # one GPU per rank: a rank's stream may spin on a flag only a kernel of ANOTHER GPU raises
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from kidney_diffusion_b200 import grid

        out = _two_stage_run(grid, torch.device("cuda", rank))
        ret[rank] = dict(shas=[_sha(t) for t in out[:2]], transport=grid.LAST_RUN.get("transport"), owned=out[2])
    finally:
        dist.destroy_process_group()


def _two_stage_run(grid, device=torch.device("cuda", 0)):
    """Stages 1 + 2 of a 4 x 4 grid as ONE pipelined plan, then the device stitch of the 256^2 results is skipped (patches are
    256^2) -- returns (stage-2 patches gathered on every rank via the stitch of their 1024^2 nearest upsample, owned count)."""
    grid._MODEL_CACHE.clear()
    grid.MODEL_PROVIDER = _provider((3, 2, 2))
    grid.CANVAS_FN = grid.default_canvas
    args = _args(inpaint_resample=2, max_batch=2, device=str(device))
    zoomed = torch.rand(1, 3, 420, 420, generator=torch.Generator().manual_seed(0)).to(device)
    cond, pos, n = grid.get_cond_images(args, zoomed, 1, lazy=True)
    o = grid.choose_orientation(pos)
    med = grid._run(1, (1, 2), args, None, cond, pos, 0.25, o, n)
    owned = sum(p is not None for p in med)
    up = grid.PatchSet([None if p is None else torch.nn.functional.interpolate(p, 1024, mode="nearest") for p in med], owner=med.owner)
    full = grid.stitch_device(zoomed, up, pos, n, 0.25, device)
    return full, full[:, :, ::4, ::4].contiguous(), owned


def test_peer_mailbox_two_ranks_equal_single_rank(cuda_lib):
    """Two ranks on two GPUs run the pipelined 2-stage plan through the CUDA-IPC mailbox and the peer stitch; every rank ends
    with the same image as the single-process run, bit for bit.  Needs 2 GPUs: ranks that wait on each other's flags must never
    share a GPU (nothing guarantees their kernels run at the same time); on a 1-GPU box the multi-rank logic is covered by the
    gloo tests on the CPU and by bench.py's `identical_to_1gpu` check at N > 1."""
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (one rank per GPU)")

    from kidney_diffusion_b200 import grid

    single = _two_stage_run(grid)
    assert single[2] == 16
    ret = mp.Manager().dict()
    mp.spawn(_two_rank_worker, args=(2, 29500 + (os.getpid() % 2000) + 31, ret), nprocs=2, join=True)
    for r in (0, 1):
        assert ret[r]["shas"][:2] == [_sha(single[0]), _sha(single[1])], f"rank {r} image differs from the single-rank run"
        assert "mailbox" in ret[r]["transport"], ret[r]["transport"]
    assert ret[0]["owned"] > 0 and ret[1]["owned"] > 0 and ret[0]["owned"] + ret[1]["owned"] == 16

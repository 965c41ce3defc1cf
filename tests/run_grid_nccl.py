"""Multi-GPU check of the patch-grid sampler (run under torchrun on N GPUs of one node, NCCL):
the N-rank run (border strips exchanged GPU-to-GPU) must reproduce the single-GPU run bit for bit.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_grid_nccl.py
"""
import os
import sys
import time
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from test_grid_gpu import _provider

    from kidney_diffusion_b200 import grid

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grid.MODEL_PROVIDER = _provider((3, 2, 2))
    args = types.SimpleNamespace(version="v_param", overlap=0.25, inpaint_resample=2, ignore_unet_1=False, num_gpus=world, device=None,
                                 max_batch=2)
    zoomed = torch.rand(1, 3, 540, 540, generator=torch.Generator().manual_seed(0))
    cond, pos, n = grid.get_cond_images(args, zoomed, 1)
    o = grid.choose_orientation(pos)

    def run():
        torch.cuda.synchronize()
        t0 = time.time()
        low = grid.generate_image_with_unet(1, 1, args, None, cond, pos, 0.25, o, n)
        med = grid.generate_image_with_unet(1, 2, args, low, cond, pos, 0.25, o, n)
        full = grid.gather_patches(med)
        torch.cuda.synchronize()
        return full, time.time() - t0, sum(p is not None for p in med)

    run()  # warm-up (graph capture, allocator)
    full_n, t_n, owned = run()
    print(f"rank {rank}: owns {owned}/{len(pos)} patches, {world}-rank run {t_n:.2f} s", flush=True)
    dist.barrier()
    if rank == 0:
        grid.DISABLE_DIST = True
        run()
        full_1, t_1, _ = run()
        same = all(torch.equal(a, b) for a, b in zip(full_n, full_1))
        print(f"GRID_NCCL world={world} patches={len(pos)} identical_to_single_gpu={same} t_world={t_n:.2f}s t_single={t_1:.2f}s speedup={t_1 / t_n:.2f}")
        grid.DISABLE_DIST = False
        assert same
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU restatement of the reference's inpaint-canvas construction (sample_ultra_res.py:149-170) for tests.

TEST INFRASTRUCTURE ONLY (tests/ may import it; the product path uses the kd_border_pack CUDA kernel).  Pinned by
tests/golden/geometry_golden.json, which holds outputs of the reference's own generate_image_distributed run on the CPU.
"""
import torch


def canvas_from_strips(S, overlap_pos, orientation, above, side, corner, device="cpu"):
    """above / side / corner: None or (strip_view, channel_stride, row_stride) as in ops.border_pack; strips are
    above [3,ov,S] (neighbour's bottom rows), side [3,S,ov] (facing columns), corner [3,ov,ov]."""
    ov = overlap_pos
    inpaint_patch = torch.zeros(3, S, S)
    inpaint_mask = torch.zeros(S, S)
    if above is not None:
        inpaint_patch[:, :ov, :] = above[0].cpu()          # :157  inpaint_patch[:, :ov, :] = above_patch[:, -ov:, :]
        inpaint_mask[:ov, :] = 1
    if side is not None:
        if orientation == -1:
            inpaint_patch[:, :, :ov] = side[0].cpu()       # :161  next_to_patch[:, :, -ov:]
            inpaint_mask[:, :ov] = 1
        else:
            inpaint_patch[:, :, -ov:] = side[0].cpu()      # :164  next_to_patch[:, :, :ov]
            inpaint_mask[:, -ov:] = 1
    if corner is not None:                                 # :166-170, no mask update
        if orientation == -1:
            inpaint_patch[:, :ov, :ov] = corner[0].cpu()
        else:
            inpaint_patch[:, :ov, -ov:] = corner[0].cpu()
    return inpaint_patch.to(device), inpaint_mask.to(torch.uint8).to(device)

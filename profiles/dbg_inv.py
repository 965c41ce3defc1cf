import os, sys, types, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from test_grid_gpu import _provider
from kidney_diffusion_b200 import grid, ops
impl = int(os.environ.get("IMPL", "0")); ops.FUSED_STATS = os.environ.get("FUSED", "1") == "1"
ops.set_conv_impl(impl)
grid.MODEL_PROVIDER = _provider((3, 2, 2)); grid.CANVAS_FN = grid.default_canvas
zoomed = torch.rand(1, 3, 420, 420, generator=torch.Generator().manual_seed(0))
outs = {}
for mb in (1, 4):
    grid._MODEL_CACHE.clear() if False else None
    args = types.SimpleNamespace(version="v_param", overlap=0.25, inpaint_resample=2, ignore_unet_1=False, num_gpus=1, device="cuda:0", max_batch=mb)
    cond, pos, n = grid.get_cond_images(args, zoomed, 1)
    o = grid.choose_orientation(pos)
    low = grid.generate_image_with_unet(1, 1, args, None, cond, pos, 0.25, o, n)
    med = grid.generate_image_with_unet(1, 2, args, low, cond, pos, 0.25, o, n)
    outs[mb] = (torch.cat(list(low)), torch.cat(list(med)))
for name, a, b in zip(("low", "med"), outs[1], outs[4]):
    print(f"impl={impl} fused={ops.FUSED_STATS} {name}: equal={torch.equal(a, b)} maxdiff={float((a - b).abs().max()):.3e} first bad patch={[int(i) for i in torch.nonzero((a - b).flatten(1).abs().amax(1) > 0).flatten()[:5]]}")

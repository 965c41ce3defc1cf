import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from helpers import U3_KW, make_pair
from kidney_diffusion_b200 import ops
orig = ops.stats_of
def checked(x):
    st = getattr(x, "_kd_stats", None)
    if st is not None and not getattr(st, "_checked", False):
        alone = ops.oct_stats(x).reduced().double().sum(1)
        fused = st.reduced().double().sum(1)
        err = float(((alone - fused).abs() / (alone.abs() + 1)).max())
        st._checked = True
        if err > 1e-4:
            bad = torch.nonzero(((alone - fused).abs() / (alone.abs() + 1)).amax(-1) > 1e-4)
            print(f"BAD fused stats: shape {tuple(x.shape)} rpt={st.rpt} tiles={st.tiles} TB={st.TB} rows={st.partial.shape[0]} err={err:.2e} bad(b,oct)={bad[:6].tolist()} n_bad={len(bad)}")
        else:
            print(f"ok  fused stats: shape {tuple(x.shape)} TB={st.TB} err={err:.1e}")
    return orig(x)
ops.stats_of = checked
import kidney_diffusion_b200.unet_exec as ue
ou, pu = make_pair(U3_KW, lowres_cond=True, seed=3)
g = torch.Generator().manual_seed(5)
B, S = 2, 128
x = torch.randn(B, 3, S, S, generator=g); lr = torch.randn(B, 3, S, S, generator=g); cond = torch.rand(B, 3, 256, 256, generator=g)
ex = pu.executor()
ex.set_conditioning(cond_images=cond.cuda(), lowres_cond_img=lr.cuda(), text_embeds=None, text_mask=None, cond_drop_prob=0.0, image_size=S)
out = ex.forward(x.cuda(), torch.tensor([2.18, -0.5]).cuda(), torch.full((B,), 0.7093).cuda())
torch.cuda.synchronize()

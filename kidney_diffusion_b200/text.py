"""Text (clinical-vector) conditioning tower of the mask-conditioned models (train.py:28-95; used by sample_cond.py:37-48
with text_embeds = [0.0, 0.5, 0.2] per sample).  Follows imagen-pytorch 1.18.5 ``Unet.forward`` steps "text conditioning":
text_to_cond -> pad to max_text_len -> null-embedding substitution -> PerceiverResampler (attn_pool) -> mean-pool ->
to_text_non_attn_cond -> added to the time conditioning; the 36 pooled tokens join the time tokens as attention context.

The tower depends on neither x nor t: it runs once per sample() call.  Linears / LayerNorms / attention are the library's
kernels (kd_linear_small, kd_layernorm_f32, kd_attn_small_f32, kd_axpby); padding, where(), cat and mean of the <= 300 token
arrays are torch indexing glue.  Its last linear is folded into the executor's time-conditioning launch:
t = [time_hiddens | lowres_hiddens | silu(h_text)] @ [Wc | Wlc | W_text2]^T + (bc + blc + b_text2).
"""
from __future__ import annotations

import torch

from . import ops
from .modules import exists


def _perceiver(ap, x):
    """PerceiverResampler.forward (depth 2, 32 latents + 4 mean-pooled latents)."""
    B, n, cd = x.shape
    xp = x + ap.pos_emb.weight.detach()[:n]
    latents = ap.latents.detach()[None].expand(B, -1, -1)
    if exists(ap.to_latents_from_mean_pooled_seq):
        ln, lin = ap.to_latents_from_mean_pooled_seq[0], ap.to_latents_from_mean_pooled_seq[1]
        mp = ops.layernorm_f32(x.mean(dim=1).contiguous(), ln.g.detach().reshape(-1).contiguous())
        mp = ops.linear_small(mp, lin.weight.detach(), lin.bias.detach()).view(B, ap.num_latents_mean_pooled, cd)
        latents = torch.cat((mp, latents), dim=-2)
    latents = latents.contiguous()
    nl = latents.shape[1]
    for attn, ff in ap.layers:
        xn = ops.layernorm_f32(xp.contiguous(), attn.norm.weight.detach(), attn.norm.bias.detach())
        lnl = ops.layernorm_f32(latents, attn.norm_latents.weight.detach(), attn.norm_latents.bias.detach())
        q = ops.linear_small(lnl.view(B * nl, cd), attn.to_q.weight.detach()).view(B, nl, -1)
        kv_in = torch.cat((xn, lnl), dim=-2).contiguous()
        kv = ops.linear_small(kv_in.view(-1, cd), attn.to_kv.weight.detach()).view(B, n + nl, -1)
        o = ops.attn_small_f32(q, kv, attn.heads, attn.scale)
        o = ops.linear_small(o.view(B * nl, -1), attn.to_out[0].weight.detach())
        o = ops.layernorm_f32(o, attn.to_out[1].weight.detach(), attn.to_out[1].bias.detach()).view(B, nl, cd)
        latents = ops.axpby(o, latents, 1.0, 1.0)
        f = ops.layernorm_f32(latents, ff[0].g.detach().reshape(-1).contiguous())
        f = ops.linear_small(f.view(B * nl, cd), ff[1].weight.detach(), post_act=ops.ACT_GELU)
        f = ops.layernorm_f32(f, ff[3].g.detach().reshape(-1).contiguous())
        f = ops.linear_small(f, ff[4].weight.detach()).view(B, nl, cd)
        latents = ops.axpby(f, latents, 1.0, 1.0)
    return latents


def text_conditioning(ex, text_embeds, text_mask, cond_drop_prob):
    u = ex.u
    if 0.0 < cond_drop_prob < 1.0:
        raise NotImplementedError("random conditioning dropout is a training-time feature; sampling uses cond_drop_prob 0 or 1")
    keep = cond_drop_prob < 1.0
    text_embeds = text_embeds.float().contiguous()
    B, L, E = text_embeds.shape
    cd, Tc, ML = ex.cd, ex.Tc, u.max_text_len
    tok = ops.linear_small(text_embeds.view(B * L, E), u.text_to_cond.weight.detach(), u.text_to_cond.bias.detach()).view(B, L, cd)[:, :ML]
    L = tok.shape[1]
    padded = torch.zeros((B, ML, cd), device=tok.device, dtype=torch.float32)
    padded[:, :L] = tok
    mask = torch.zeros((B, ML), device=tok.device, dtype=torch.bool)
    if exists(text_mask):
        mask[:, :L] = text_mask[:, :ML].bool()
    else:
        mask[:, :L] = True
    keep_embed = mask[:, :, None] & keep
    tokens = torch.where(keep_embed, padded, u.null_text_embed.detach())
    if exists(u.attn_pool):
        tokens = _perceiver(u.attn_pool, tokens.contiguous())
    tokens = tokens.contiguous()
    mean_pooled = tokens.mean(dim=-2).contiguous()
    net = u.to_text_non_attn_cond
    h = ops.layernorm_f32(mean_pooled, net[0].weight.detach(), net[0].bias.detach())
    hidden_in = ops.linear_small(h, net[1].weight.detach(), net[1].bias.detach(), post_act=ops.ACT_SILU)  # silu(Linear(LN(mean)))
    if keep:
        tc_w = torch.cat((ex.tc_w, net[3].weight.detach()), dim=1).contiguous()
        tc_b = (ex.tc_b + net[3].bias.detach()).contiguous()
    else:  # text_hiddens replaced by null_text_hidden
        tc_w = torch.cat((ex.tc_w, torch.zeros_like(net[3].weight)), dim=1).contiguous()
        tc_b = (ex.tc_b + u.null_text_hidden.detach().reshape(-1)).contiguous()
    return dict(tokens=tokens, hidden_in=hidden_in.contiguous(), tc_w=tc_w, tc_b=tc_b)

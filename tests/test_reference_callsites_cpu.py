"""The drop-in boundary against the REFERENCE'S OWN code: the factory functions of the reference's training
scripts (`unet_generator`, `FixedNullUnet`, `init_imagen`) are extracted from /root/reference with `ast` -- nothing is copied into
the repo -- and executed with `imagen_pytorch` resolving to this repo's shim.  The models they construct must equal, key by key
and shape by shape, what kidney_diffusion_b200.factories builds from its restated argument tables; the scripts' own module-level
imports (datasets needing slideio / h5py / wandb ...) are never executed.

Runs only where /root/reference exists (the build container); the GPU box does not have it.
"""
import ast
import os

import pytest
import torch
from torch import nn

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present on this machine")

WANTED_FUNCS = {"unet_generator", "init_imagen"}
WANTED_CLASSES = {"FixedNullUnet"}


def load_reference_factories(script):
    """Namespace holding the reference script's factory definitions (and its ALL-CAPS module constants), with the names they
    use (`Unet`, `Imagen`, `NullUnet`, `nn`, `torch`) bound to the shim."""
    import imagen_pytorch as shim

    assert "kidney_diffusion_b200" in shim.Unet.__module__, "imagen_pytorch must resolve to the repo's shim"
    path = os.path.join(REF, script)
    tree = ast.parse(open(path).read(), filename=path)
    keep = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED_FUNCS:
            keep.append(node)
        elif isinstance(node, ast.ClassDef) and node.name in WANTED_CLASSES:
            keep.append(node)
        elif isinstance(node, ast.Assign) and all(isinstance(t, ast.Name) and t.id.isupper() for t in node.targets):
            keep.append(node)
    ns = dict(Unet=shim.Unet, Imagen=shim.Imagen, NullUnet=shim.NullUnet, ImagenTrainer=shim.ImagenTrainer, SRUnet1024=shim.SRUnet1024,
              ElucidatedImagen=shim.ElucidatedImagen, torch=torch, nn=nn)
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    assert {"unet_generator", "init_imagen", "FixedNullUnet"} <= set(ns), f"{script}: factory definitions not found"
    return ns


def describe(imagen):
    sd = {k: tuple(v.shape) for k, v in imagen.state_dict().items()}
    meta = dict(image_sizes=tuple(imagen.image_sizes), objectives=tuple(imagen.pred_objectives),
                timesteps=tuple(s.num_timesteps for s in imagen.noise_schedulers),
                schedules=tuple(s.noise_schedule for s in imagen.noise_schedulers), lowres=tuple(bool(u.lowres_cond) for u in imagen.unets),
                condition_on_text=imagen.condition_on_text, text_embed_dim=imagen.text_embed_dim, random_crop_sizes=tuple(imagen.random_crop_sizes),
                unet_types=tuple("Unet" if hasattr(u, "init_conv") else "Null" for u in imagen.unets))
    return sd, meta


CASES = [
    # script, kwargs style, our factory
    ("train_ultra_res_v_param.py", "mag", dict(version="v_param")),
    ("train_ultra_res.py", "mag", dict(version="")),
    ("train_ultra_res_v2.py", "mag", dict(version="v2")),
    ("train_ultra_res_airs.py", "mag", dict(version="airs")),
    ("train_uncond.py", "plain", "uncond"),
    ("train.py", "plain", "cond"),
]


@pytest.mark.parametrize("script,style,ours", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("unet_number", [1, 2, 3])
def test_reference_init_imagen_builds_the_same_model_as_factories(script, style, ours, unet_number, monkeypatch):
    from kidney_diffusion_b200 import factories
    from kidney_diffusion_b200.imagen import Imagen

    if not os.path.exists(os.path.join(REF, script)):
        pytest.skip(f"{script} not in the reference checkout")
    monkeypatch.setattr(Imagen, "cuda", lambda self, *a, **k: self)  # train.py / train_uncond.py end with .cuda()
    ns = load_reference_factories(script)
    with torch.device("meta"):  # shapes and names only: no 2.7 GB of parameters on the host
        for mag in ((0, 1) if style == "mag" else (None,)):
            if style == "mag":
                theirs = ns["init_imagen"](mag, unet_number, device=torch.device("meta"))
                mine = factories.init_imagen_ultra_res(mag, unet_number, **ours)
            elif ours == "uncond":
                theirs = ns["init_imagen"](unet_number)
                mine = factories.init_imagen_uncond(unet_number)
            else:
                theirs = ns["init_imagen"](unet_number)
                mine = factories.init_imagen_cond(unet_number)
            sd_t, meta_t = describe(theirs)
            sd_m, meta_m = describe(mine)
            assert meta_t == meta_m, (script, mag, unet_number, meta_t, meta_m)
            assert list(sd_t) == list(sd_m), f"{script} unet {unet_number}: state-dict key order differs"
            assert sd_t == sd_m
            assert len(sd_t) > (100 if unet_number else 0)
            # the reference's own FixedNullUnet subclass works against the shim's NullUnet
            nulls = [u for u in theirs.unets if isinstance(u, ns["FixedNullUnet"])]
            assert len(nulls) == 2 and all(hasattr(u, "dummy_parameter") for u in nulls)


def test_reference_sampling_call_sites_use_only_supported_arguments():
    """Every keyword the reference passes to `imagen.sample` / `trainer.sample` / `trainer.load` / `Imagen(...)` / `Unet(...)` in its
    sampling scripts exists in the shim's signatures (checked on the reference's AST)."""
    import inspect

    import imagen_pytorch as shim

    sample_params = set(inspect.signature(shim.Imagen.sample).parameters)
    trainer_sample = sample_params | set(inspect.signature(shim.ImagenTrainer.sample).parameters)
    seen = 0
    for script in ("sample_ultra_res.py", "sample_cond.py", "sample_uncond.py", "sample.py", "outpainting.py"):
        path = os.path.join(REF, script)
        if not os.path.exists(path):
            continue
        for node in ast.walk(ast.parse(open(path).read())):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "sample":
                owner = ast.unparse(node.func.value)
                kws = {k.arg for k in node.keywords if k.arg}
                allowed = trainer_sample if "trainer" in owner else sample_params
                assert kws <= allowed, f"{script}: {owner}.sample(...) passes unsupported {sorted(kws - allowed)}"
                seen += 1
    assert seen >= 4

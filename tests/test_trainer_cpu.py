"""ImagenTrainer.load semantics on the host (reference use: sample_cond.py:26-35, sample_uncond.py:23-33; scope rows a16 / N1):
strict load, restore_parts fallback, and the EMA fold -- trainer.sample() of imagen-pytorch runs the EMA copies of the unets,
stored in the checkpoint under 'ema' as '<unet index>.ema_model.<key>'."""
import torch

from kidney_diffusion_b200 import Imagen, ImagenTrainer, Unet, restore_parts
from kidney_diffusion_b200.factories import FixedNullUnet


def tiny(seed=0):
    torch.manual_seed(seed)
    return Imagen(unets=(Unet(dim=64, dim_mults=(1, 2), layer_attns=False, layer_cross_attns=False), FixedNullUnet(lowres_cond=True)),
                  image_sizes=(16, 64), timesteps=(2, 2), condition_on_text=False)


def test_trainer_load_folds_ema_weights(tmp_path):
    src = tiny(1)
    ema = {f"0.ema_model.{k}": v + 0.5 for k, v in src.unets[0].state_dict().items()}
    ema["0.initted"] = torch.tensor(True)
    ema["0.step"] = torch.tensor(10)
    path = tmp_path / "ckpt.pt"
    torch.save(dict(model=src.state_dict(), ema=ema, version="1.18.5"), path)
    dst = tiny(2)
    trainer = ImagenTrainer(imagen=dst)
    trainer.load(str(path))
    for k, v in src.unets[0].state_dict().items():
        assert torch.equal(dst.unets[0].state_dict()[k], v + 0.5), f"{k}: trainer.sample() must run the EMA weights"
    # use_ema=False keeps the online weights
    dst2 = tiny(3)
    ImagenTrainer(imagen=dst2, use_ema=False).load(str(path))
    for k, v in src.state_dict().items():
        assert torch.equal(dst2.state_dict()[k], v)


def test_trainer_load_partial_restore_and_noop(tmp_path, capsys):
    src = tiny(1)
    sd = dict(src.state_dict())
    victim = next(k for k in sd if k.endswith("final_conv.bias"))
    sd[victim] = torch.zeros(7)  # wrong shape -> strict load fails -> restore_parts copies everything else
    sd["unets.0.extra.weight"] = torch.zeros(3)
    path = tmp_path / "ckpt.pt"
    torch.save(dict(model=sd, version="1.17.0"), path)
    dst = tiny(2)
    before = dst.state_dict()[victim].clone()
    ImagenTrainer(imagen=dst).load(str(path))
    out = capsys.readouterr().out
    assert "Trying partial load" in out and "1.17.0" in out
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], before if k == victim else v)
    assert ImagenTrainer(imagen=dst).load(str(tmp_path / "missing.pt"), noop_if_not_exist=True) is None
    # restore_parts on plain dicts (sample_ultra_res.py:63)
    tgt = {"a": torch.zeros(2), "b": torch.zeros(3)}
    restore_parts(tgt, {"a": torch.ones(2), "b": torch.ones(4), "c": torch.ones(1)})
    assert torch.equal(tgt["a"], torch.ones(2)) and torch.equal(tgt["b"], torch.zeros(3))

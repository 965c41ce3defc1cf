from kidney_diffusion_b200.trainer import ImagenTrainer, restore_parts  # noqa: F401

from kidney_diffusion_b200.trainer import __version__  # noqa: F401

"""damc_b200 -- B200-native sampling hot path of Diffusion-Amortized MCMC.

``from damc_b200.MCMC import sample_langevin_prior_z, sample_langevin_post_z_with_prior`` mirrors the reference's
``from src.MCMC import ...``; ``damc_b200.diffusion_net`` mirrors ``src.diffusion_net``."""
__all__ = ["MCMC", "diffusion_net", "parallel"]

"""lsnf_b200 -- B200-native short-run Langevin posterior inference for latent-space normalizing-flow priors."""

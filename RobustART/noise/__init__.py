from RobustART.noise.add_noise import AddNoise  # noqa: F401

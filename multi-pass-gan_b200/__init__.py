"""multi-pass-gan_b200: B200-native (sm_100a) generator hot path of maxwerhahn/Multi-pass-GAN.

The directory name is not a valid Python identifier; import it as ``mpgan_b200`` (the shim
``mpgan_b200.py`` at the repository root registers this package under that name).
"""
__version__ = "0.1.0"

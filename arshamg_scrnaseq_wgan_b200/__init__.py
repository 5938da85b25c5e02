"""B200-native WGAN(-GP) hot path of luslab/arshamg-scrnaseq-wgan (scripts/WGAN-GP_minimal.py).

Python host layer over the sm_100a CUDA library (csrc/, C ABI in include/wgan_b200.h).
"""
from ._lib import build, load, WganError  # noqa: F401
from .trainer import (WGANConfig, WGANTrainer, Adam_Prediction_Optimizer, generator, discriminator,  # noqa: F401
                      PARAM_NAMES, GEN_NAMES, CRITIC_NAMES, NET_GENERATOR, NET_CRITIC)

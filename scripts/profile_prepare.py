"""A few wgan_critic_prepare calls (stage z, generator forward with the interpolate fused into the output product's
epilogue): target of `ncu --set full -k regex:gemm_tf32_2cta -s 17 -c 1` = the fused-interpolate product of the 6th call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
B = 4096
tr = WGANTrainer(WGANConfig(), max_batch=B, device=0)
tr.init_params(seed=1)
x = tr.real_slot(B); x.uniform_()
z = tr.noise_prior(B, seed=5, call_id=1, row0=0)
eps = tr.uniform(B, seed=5, call_id=1, row0=0)
for _ in range(8):
    tr.critic_prepare(x, z, eps)
torch.cuda.synchronize()

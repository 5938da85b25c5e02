"""A few critic updates (target of `ncu --set full -k regex:<kernel>`): python scripts/profile_step.py [explicit|gram]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
form = sys.argv[1] if len(sys.argv) > 1 else "explicit"
B = 4096
tr = WGANTrainer(WGANConfig(gp_form=form), max_batch=B, device=0)
tr.init_params(seed=1)
g = torch.Generator(device="cuda:0").manual_seed(0)
x = torch.rand((B, tr.cfg.gex_size), device="cuda:0", generator=g)
z = tr.noise_prior(B, seed=5, call_id=1, row0=0)
eps = tr.uniform(B, seed=5, call_id=1, row0=0)
for _ in range(4):
    tr.critic_step(x, z, eps, read=False)
torch.cuda.synchronize()

"""Timing of the other BASELINE configs on one B200 (informational; bench.py reports config 2 only):
config 3 = critic + gradient-penalty updates at batch 65536 (reference dims), config 4 = widened networks
(4x hidden width, 20000 genes) at batch 4096."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer

def time_fn(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# config 3: critic-only steps with GP at batch 65536 (30.70 MFLOP/sample)
B = 65536
cfg = WGANConfig()
tr = WGANTrainer(cfg, max_batch=B, device=0); tr.init_params(0)
x = tr.real_slot(B); x.uniform_()
z = tr.noise_prior(B, seed=1, call_id=0, row0=0); eps = tr.uniform(B, 1, 0, row0=0)
ms = time_fn(lambda: tr.critic_step(x, z, eps, read=False), 10)
print(f"config 3 (1 GPU): critic+GP update at batch {B}: {ms:.2f} ms = {1000/ms:.1f} updates/s, "
      f"{B*3.070e7/ms/1e9:.0f} TFLOP/s algorithmic, {B/ms*1e3/1e6:.1f} M samples/s", flush=True)
tr.close(); del tr, x, z, eps
torch.cuda.empty_cache()

# config 4: wide networks, batch 4096 (critic step 376.8 MFLOP/sample, generator step 390.1, iteration 2.274 GFLOP/sample)
dims = dict(noise_input_size=100, inflate_to_size=2400, gex_size=20000, disc_internal_size=800)
B = 4096
cfg = WGANConfig(**dims)
tr = WGANTrainer(cfg, max_batch=B, device=0); tr.init_params(0)
tr.set_data(torch.rand((1500, 20000), device="cuda:0"))
replay = tr.capture_iteration(B)
ms = time_fn(replay, 10)
print(f"config 4 (1 GPU): wide WGAN-GP iteration at batch {B}: {ms:.2f} ms = {1000/ms:.1f} iters/s, "
      f"{B*2.274e9/ms/1e9:.0f} TFLOP/s algorithmic", flush=True)
zs = tr.noise_prior(B, seed=2, call_id=9)
out = torch.empty((B, tr.Gl), device="cuda:0")[:, :20000]
ms = time_fn(lambda: tr.generator(zs, out=out), 10)
print(f"config 4 sampler: {B/ms*1e3/1e6:.2f} M cells/s ({B*1.080e8/ms/1e9:.0f} TFLOP/s)")

"""Secondary measurements quoted in DESIGN.md: the reference's own configuration (batch 32, no penalty,
generator-then-critic [W:11,196-199]) on the B200 vs the CPU restatement, and the latent-fit step rate."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
from bench import synthetic_cells

def timed_replay(replay, n):
    for _ in range(5):
        replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

data = synthetic_cells(1500, 6605, 0, "cuda:0")
for name, cfg, B in (("reference-literal (batch 32, lambda 0, 1 gen + 1 critic update)", WGANConfig.reference_literal(), 32),
                     ("WGAN-GP (batch 32, lambda 10, 5 critic + 1 gen)", WGANConfig(), 32),
                     ("WGAN-GP (batch 512)", WGANConfig(), 512)):
    tr = WGANTrainer(cfg, max_batch=B, device=0)
    tr.init_params(0); tr.set_data(data)
    ms = timed_replay(tr.capture_iteration(B), 200)
    print(f"B200 {name}: {ms:.3f} ms/iteration = {1000/ms:.0f} iters/s (CUDA graph, {tr.graph_launches_per_replay} launches)", flush=True)
    tr.close()

# CPU restatement at the reference's own configuration
from oracle import wgan_oracle as O
from oracle.wgan_torch_cpu import TorchCpuTrainer
torch.set_num_threads(len(os.sched_getaffinity(0)))
ocfg = O.OracleConfig()
t = TorchCpuTrainer(ocfg, O.glorot_init(ocfg, 0, dtype=np.float32))
dcpu = data.cpu()
def cpu_it():
    idx = torch.randint(1500, (32,))
    z = (torch.randn(32, 100) * 0.1 + torch.poisson(torch.ones(32, 100))).abs()
    t.reference_iteration(dcpu[idx], z)
for _ in range(5): cpu_it()
t0 = time.perf_counter()
for _ in range(50): cpu_it()
dt = (time.perf_counter() - t0) / 50
print(f"CPU restatement reference-literal batch 32: {dt*1e3:.2f} ms/iteration = {1/dt:.1f} iters/s on {torch.get_num_threads()} threads", flush=True)

# latent fit (config 5): 65536 targets, one TF-Adam step on z
B = 32768
tr = WGANTrainer(WGANConfig(), max_batch=B, device=0)
tr.init_params(0)
target = tr.sample_cells(B, seed=3)
z, loss = tr.latent_fit(target, steps=3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
z, loss = tr.latent_fit(target, steps=20, z0=z)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"latent fit: {ms:.3f} ms per Adam step on {B} cells = {B/ms*1e3/1e6:.1f} M cell-steps/s, loss {float(loss.mean()):.4f}")

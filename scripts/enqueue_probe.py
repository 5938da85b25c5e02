"""How long does the host take to ENQUEUE one iteration (no GPU wait) vs the GPU time per iteration?"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
sys.path.insert(0, ROOT)
from bench import synthetic_cells
for B in (4096, 512, 32):
    tr = WGANTrainer(WGANConfig(), max_batch=B, device=0)
    tr.init_params(0)
    tr.set_data(synthetic_cells(1500, 6605, 0, "cuda:0"))
    for it in range(5):
        tr.train_iteration_resident(it, B)
    torch.cuda.synchronize()
    n = 20
    t0 = time.perf_counter()
    for it in range(n):
        tr.train_iteration_resident(5 + it, B)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"B={B}: host enqueue {1e3*(t1-t0)/n:.3f} ms/iter, total {1e3*(t2-t0)/n:.3f} ms/iter", flush=True)
    tr.close()

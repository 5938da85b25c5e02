"""Times wgan_allreduce_grads alone (both arenas) under torchrun: python -m torch.distributed.run ... allreduce_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
from arshamg_scrnaseq_wgan_b200.trainer import NET_CRITIC, NET_GENERATOR, KIND_GRAD
tr = WGANTrainer(WGANConfig(lambda_gp=10.0), max_batch=64, device=local)
for net, name in ((NET_CRITIC, "critic"), (NET_GENERATOR, "generator")):
    a = tr.arena(net, KIND_GRAD)
    a.normal_()
    for backend in ("engine", "nccl"):
        tr.dp_backend = backend
        for _ in range(5):
            tr._allreduce(net)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            tr._allreduce(net)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        if dist.get_rank() == 0:
            mb = a.numel() * 4 / 1e6
            w = dist.get_world_size()
            print(f"{name:9s} {mb:6.2f} MB  {backend:6s} {us:7.1f} us   bus {2 * (w - 1) / w * mb / us * 1e3:7.1f} GB/s", flush=True)
dist.barrier()
os._exit(0)

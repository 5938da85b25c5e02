"""Kernel timeline of graph-replayed data-parallel iterations (CUPTI through torch.profiler): where each gradient
exchange starts / ends relative to the kernels around it.  torchrun --nproc-per-node N scripts/probes/dp_timeline.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
local = int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
B = 4096
tr = WGANTrainer(WGANConfig(lambda_gp=10.0, n_critic=5), max_batch=B, device=local)
tr.init_params(seed=0)
g = torch.Generator(device=f"cuda:{local}").manual_seed(0)
tr.set_data(torch.rand((1500, tr.cfg.gex_size), device=f"cuda:{local}", generator=g))
for it in range(2):
    tr.train_iteration_resident(it, B, seed=0)
replay = tr.capture_iteration(B, seed=0, start_it=2)
for _ in range(5):
    replay()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        replay()
    torch.cuda.synchronize()
ev = [e for e in json.loads(open(prof.export_chrome_trace(f"/tmp/trace{local}.json") or f"/tmp/trace{local}.json").read())["traceEvents"]
      if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
out = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "gpurun_out",
                   f"dp_timeline_w{world}_rank{local}.txt")
with open(out, "w") as f:
    prev_end = t0
    for e in ev:
        name = e["name"].split("(")[0].replace("void ", "")[:60]
        f.write(f"{e['ts'] - t0:10.1f} {e['dur']:8.1f} end {e['ts'] - t0 + e['dur']:10.1f} stream {e['args'].get('stream')} {name}\n")
if world > 1:
    dist.barrier()
os._exit(0)

#!/usr/bin/env python
"""Benchmark of the WGAN-GP hot path (BASELINE.json metric: train iters/sec and generated cells/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2|3|4|5]

`--config` selects the BASELINE.json configuration (default 2, the one the headline metric is quoted on):

  2  one WGAN-GP training iteration = n_critic 5 critic updates with gradient penalty (lambda 10) + 1 generator
     update, reference architecture (100-600-600-6605 generator, 6605-200-200-1 critic), batch 4096 PER GPU
     (weak scaling), TF-Adam, synthetic lTPM-like data of the CSV's gene dimension, 1500 resident cells [W:18].
  3  gradient-penalty stress: the n_critic loop alone (5 critic + penalty updates per step) at a FIXED global batch
     of 65536, sharded over the N GPUs (strong scaling: local batch 65536 / N).
  4  config 2 with widened networks (4x hidden width: 100-2400-2400-20000 / 20000-800-800-1), batch 4096 per GPU.
  5  generator-only sampling of 10 M cells [W:217-221], the cell range sharded over the N GPUs without
     communication (strong scaling), plus the latent-vector TF-Adam fit as an extra key.

For N > 1 the driver launches one rank per GPU with torchrun; gradients are summed once per optimizer step by the
library's own exchange kernel over NVLink peer memory (wgan_allreduce_grads / wgan_generator_grads_exchange).
`value`, `e2e` and `roofline` refer to the explicit gradient-penalty chain (the work the FLOP table of SURVEY 8d
counts); the Gram form of the penalty is timed as well and reported SEPARATELY under `gram_form` (config 2).

`--impl reference` times the reference's CPU implementation of the same configuration: TensorFlow 1.x is not
installable here, so this is the FP32 torch-CPU restatement of the reference graph (oracle/wgan_torch_cpu.py) on all
host cores, on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_CELLS = 1500            # num_cells_train [W:18]
REF_DIMS = dict(noise_input_size=100, inflate_to_size=600, gex_size=6605, disc_internal_size=200)
WIDE_DIMS = dict(noise_input_size=100, inflate_to_size=2400, gex_size=20000, disc_internal_size=800)
STRESS_BATCH = 65536
SAMPLE_CELLS = 10_000_000
SAMPLE_CHUNK = 32768


def flops(dims):
    """Algorithmic GEMM FLOPs per sample of one generator forward (2 per multiply-add)."""
    Z, Hg, G = dims["noise_input_size"], dims["inflate_to_size"], dims["gex_size"]
    return {"gen_fwd": 2.0 * (Z * Hg + Hg * Hg + Hg * G)}


CONFIGS = {
    2: dict(metric="wgan_gp_train_iters_per_sec", scaling="weak", dims=REF_DIMS, batch=4096,
            unit="iters/s (batch-4096 iterations: 5 critic+GP updates + 1 generator update each)",
            # 185.1 MFLOP per sample and iteration
            flop_per_sample_step=1.851e8,
            workload="BASELINE config 2: WGAN-GP iteration (n_critic 5 critic+penalty updates + 1 generator update), "
                     "reference dims 100-600-600-6605 / 6605-200-200-1, batch 4096 per GPU, lambda 10, TF-Adam lr 1e-5, "
                     "synthetic lTPM 1500x6605"),
    3: dict(metric="wgan_gp_critic_loops_per_sec", scaling="strong", dims=REF_DIMS, batch=STRESS_BATCH,
            unit="n_critic loops/s (5 critic+penalty updates each at global batch 65536)",
            flop_per_sample_step=5 * 3.070e7,
            workload="BASELINE config 3: gradient-penalty stress, the n_critic loop alone (5 critic+penalty updates per "
                     "step), reference dims, GLOBAL batch 65536 sharded over the GPUs, lambda 10, TF-Adam, "
                     "synthetic lTPM 1500x6605"),
    4: dict(metric="wgan_gp_train_iters_per_sec", scaling="weak", dims=WIDE_DIMS, batch=4096,
            unit="iters/s (batch-4096 iterations of the widened networks: 5 critic+GP updates + 1 generator update each)",
            flop_per_sample_step=2.274e9,
            workload="BASELINE config 4: WGAN-GP iteration (n_critic 5 + 1) with 4x hidden width and a full-transcriptome "
                     "gene dimension, dims 100-2400-2400-20000 / 20000-800-800-1, batch 4096 per GPU, lambda 10, TF-Adam, "
                     "synthetic lTPM 1500x20000"),
    5: dict(metric="generated_cells_per_sec", scaling="strong", dims=REF_DIMS, batch=SAMPLE_CHUNK,
            unit="cells/s (generator-only sampling, noise prior drawn on the device)",
            flop_per_sample_step=None,
            workload="BASELINE config 5: generator-only sampling of 10 M synthetic cells (noise prior + G(z)), reference "
                     "dims 100-600-600-6605, the cell range sharded over the GPUs, chunks of 32768 cells"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="override the configuration's batch (experiments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gp-form", default="explicit", choices=["explicit", "gram"],
                    help="penalty evaluation of the headline run (the Gram form is also timed as `gram_form`)")
    ap.add_argument("--no-gram", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip sampler / reference-literal / latent-fit extras")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_cells(n_cells, n_genes, seed, device):
    """log2(TPM+1)-like cells x genes matrix, per-gene min-max scaled to [0,1] as [W:34-39]
    (per-gene dropout probability and gamma magnitude)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    p_drop = torch.empty(1, n_genes).uniform_(0.3, 0.95, generator=g)
    scale = torch.empty(1, n_genes).uniform_(2.0, 12.0, generator=g)
    expressed = torch.rand(n_cells, n_genes, generator=g) > p_drop
    mag = torch._standard_gamma(torch.full((n_cells, n_genes), 2.0), generator=g) * (scale / 2.0)
    raw = torch.where(expressed, mag, torch.zeros(()))
    lo = raw.min(0, keepdim=True).values
    rng = raw.max(0, keepdim=True).values - lo
    rng[rng == 0] = 1.0
    return ((raw - lo) / rng).to(device=device, dtype=torch.float32)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: FP32 torch-CPU restatement of the reference graph
# ------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_step_runner(config, batch, seed=0, literal=False):
    """Returns (run_one_step, description) for the CPU restatement of `config` at `batch` rows."""
    import numpy as np
    import torch
    from oracle import wgan_oracle as O
    from oracle.wgan_torch_cpu import TorchCpuTrainer
    dims = CONFIGS[config]["dims"]
    cfg = O.OracleConfig(lambda_gp=0.0 if literal else 10.0, n_critic=1 if literal else 5, **dims)
    tr = TorchCpuTrainer(cfg, O.glorot_init(cfg, seed, dtype=np.float32))
    g = torch.Generator().manual_seed(seed + 1)

    def noise(b):
        return (torch.randn(b, cfg.noise_input_size, generator=g) * 0.1 +
                torch.poisson(torch.ones(b, cfg.noise_input_size), generator=g)).abs()

    if config == 5:
        def sample_chunk():
            return tr.sample(noise(batch))
        return sample_chunk, "noise prior + generator forward"
    data = synthetic_cells(N_CELLS, cfg.gex_size, seed, "cpu")

    def feeds(n):
        xs, zs, es = [], [], []
        for _ in range(n):
            idx = torch.randint(N_CELLS, (batch,), generator=g)         # [W:188]
            xs.append(data[idx]); zs.append(noise(batch)); es.append(torch.rand(batch, generator=g))
        return xs, zs, es

    if literal:
        def literal_iteration():
            xs, zs, _ = feeds(1)
            return tr.reference_iteration(xs[0], zs[0])
        return literal_iteration, "generator update then critic update on one batch [W:196,199], no penalty"
    if config == 3:
        def critic_loop():
            out = {}
            for x, z, e in zip(*feeds(cfg.n_critic)):
                out = tr.critic_step(x, z, e)[0]
            return out
        return critic_loop, "5 critic+penalty updates"

    def one_iteration():
        xs, zs, es = feeds(cfg.n_critic)
        return tr.wgan_gp_iteration(xs, zs, es, noise(batch))
    return one_iteration, "5 critic+GP updates + 1 generator update"


def time_cpu(config, batch, steps, warmup, budget_s, literal=False):
    """Times `steps` steps of the CPU restatement after `warmup`; when that would exceed budget_s a smaller batch sample
    is timed and scaled linearly in the batch (stated in `sample`).  Returns (value in the config's unit per second
    basis, ms per full step, threads, sample text)."""
    import torch
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would silently make the
    # CPU arm single-threaded)
    torch.set_num_threads(max(1, host_cores()))
    # probe at a bounded batch, then time the largest power-of-two fraction of the batch that fits the budget
    b = min(batch, 512)
    run, what = cpu_step_runner(config, b, literal=literal)
    run()
    t0 = time.perf_counter(); run(); t_probe = time.perf_counter() - t0
    per_row = t_probe / b
    while b * 2 <= batch and (steps + warmup) * per_row * (b * 2) <= budget_s:
        b *= 2
    if b * 2 > batch and (steps + warmup) * per_row * batch <= budget_s:
        b = batch
    while (steps + warmup) * per_row * b > budget_s and b > 64:
        b //= 2
    run, what = cpu_step_runner(config, b, literal=literal)
    run()
    for _ in range(max(warmup - 1, 0)):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    steps_per_s = steps / dt * (b / batch)            # full-batch steps per second
    sample = (f"{steps} timed steps ({what}) at batch {b}"
              + ("" if b == batch else f", scaled by {b}/{batch} to batch-{batch} steps")
              + f" after {warmup} warm-up; FP32 torch-CPU restatement of the reference graph (TensorFlow not installable)")
    return steps_per_s, 1e3 / steps_per_s, torch.get_num_threads(), sample


def cpu_value(config, steps_per_s, batch):
    """The config's metric from full-batch steps per second."""
    return steps_per_s * batch if config == 5 else steps_per_s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    batch = args.batch or c["batch"]
    sps, ms, cores, sample = time_cpu(args.config, batch, args.steps, args.warmup, budget_s=200.0)
    v = cpu_value(args.config, sps, batch)
    line = {"impl": "reference", "metric": c["metric"], "value": v, "unit": c["unit"], "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": c["workload"], "baseline_config": args.config},
            "cpu_baseline": {"value": v, "unit": c["unit"], "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": c["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def time_kernel(fn, iters, torch):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def measure_tf32_peak(torch):
    """Dense TF32 peak the same way MEASURED_PEAKS.json's bf16 figure is taken: cuBLAS 8192^3 with allow_tf32, best of
    a short burst (TF32 is not in that file).  Used as the denominator of tensor-bound fractions."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn((n, n), device="cuda"); b = torch.randn((n, n), device="cuda")
        best = 0.0
        for _ in range(3):
            ms = time_kernel(lambda: torch.matmul(a, b), 5, torch)
            best = max(best, 2.0 * n ** 3 / ms / 1e9)
        del a, b
        return best
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def roofline_entry(name, ms, flops_, bytes_, per_step, pk, tf32_peak):
    t_tensor = flops_ / (tf32_peak * 1e12)
    t_hbm = bytes_ / (pk["hbm_gbs"] * 1e9)
    r = {"kernel": name, "ms": ms, "tflops": flops_ / ms / 1e9, "gbs": bytes_ / ms / 1e6, "per_step": per_step,
         "flops": flops_, "bytes": bytes_}
    if t_hbm >= t_tensor:
        r.update({"bound": "hbm", "achieved": r["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                  "frac": r["gbs"] / pk["hbm_gbs"]})
    else:
        r.update({"bound": "tensor", "achieved": r["tflops"], "peak": tf32_peak, "unit": "TFLOP/s",
                  "frac": r["tflops"] / tf32_peak})
    return r


def dominant_kernel_roofline(tr, batch, torch, pk, tf32_peak, n_critic=5, generator_update=True):
    """Times the step's heaviest kernel instances in isolation (CUDA events on the launch stream, operands larger than
    L2 or re-streamed) and reports the dominant one (time x launches per step) against its binding roof."""
    import ctypes as C
    cfg = tr.cfg
    G, Gl, Hd, Hg = cfg.gex_size, tr.Gl, cfg.disc_internal_size, cfg.inflate_to_size
    dev = f"cuda:{tr.device}"
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    null = C.c_void_p()
    tiles = C.c_int(0)
    Hdl = tr.lib.wgan_padded_ld(Hd)
    # MN-major TMA operands are read in whole 32-float chunks: the last chunk of the last row may touch up to 124 B
    # past the logical end (include/wgan_b200.h, wgan_gemm), so these stand-alone operands carry one spare row
    def spare(rows, ld, sub=0.0):
        return (torch.rand((rows + 1, ld), device=dev) - sub)[:rows]
    X3 = spare(3 * batch, Gl)
    W1 = spare(G, Hdl, 0.5)
    b1 = torch.zeros((Hd,), device=dev)
    H1 = torch.empty((3 * batch, Hdl), device=dev)
    D1 = spare(3 * batch, Hdl, 0.5)
    GW1 = torch.empty((G, Hdl), device=dev)
    GH2 = torch.rand((batch, tr.lib.wgan_padded_ld(Hg)), device=dev)
    WG3 = spare(Hg, Gl, 0.5)
    bg3 = torch.zeros((G,), device=dev)
    XT = torch.empty((batch, Gl), device=dev)

    def gemm(a_mn, b_mn, A, B, D, M, N, K, bias=None, act=0):
        rc = tr.lib.wgan_gemm(tr._h, a_mn, b_mn, p(A), A.stride(0), p(B), B.stride(0), p(D), D.stride(0), M, N, K,
                              p(bias) if bias is not None else null, act, null, 0, null, null, C.byref(tiles), 0, st())
        assert rc == 0, tr.lib.wgan_last_error(tr._h)

    out = []
    ms = time_kernel(lambda: gemm(0, 1, X3, W1, H1, 3 * batch, Hd, G, b1, 1), 20, torch)
    out.append(roofline_entry(f"critic layer-1 forward [3B x {G}] x [{G} x {Hd}] on the stacked rows [fake | interpolate | "
                              "real] (TF32 tcgen05 cta_group::2, stream-K, bias+LeakyReLU epilogue)", ms,
                              2.0 * 3 * batch * G * Hd, 4.0 * (3 * batch * G + G * Hd + 3 * batch * Hd), n_critic, pk, tf32_peak))
    ms = time_kernel(lambda: gemm(1, 1, X3, D1, GW1, G, Hd, 3 * batch), 20, torch)
    out.append(roofline_entry(f"critic layer-1 weight gradient [{G} x 3B] x [3B x {Hd}] (TF32 tcgen05 cta_group::2, stream-K)",
                              ms, 2.0 * 3 * batch * G * Hd, 4.0 * (3 * batch * G + 3 * batch * Hd + G * Hd), n_critic, pk, tf32_peak))
    ms_plain = time_kernel(lambda: gemm(0, 1, GH2, WG3, XT, batch, G, Hg, bg3, 1), 20, torch)
    out.append(roofline_entry(f"generator output forward [B x {Hg}] x [{Hg} x {G}] (TF32 tcgen05 cta_group::2, bias+LeakyReLU "
                              "epilogue): generator update and sampler", ms_plain, 2.0 * batch * G * Hg,
                              4.0 * (batch * Hg + Hg * G + batch * G), 1 if generator_update else 0, pk, tf32_peak))
    # the same product as it runs inside a critic update: the interpolate x^ = eps x + (1 - eps) x~ is a second
    # output of its epilogue (reads x, writes x~ and x^).  Timed in situ as wgan_critic_prepare - wgan_generator_prepare
    # (identical launches except for that epilogue) on top of the plain product.
    x = tr.real_slot(batch); x.uniform_()
    z = tr.noise_prior(batch, seed=3, call_id=5)
    eps = tr.uniform(batch, seed=3, call_id=5)
    t_cp = time_kernel(lambda: tr.critic_prepare(x, z, eps), 20, torch)
    t_gp = time_kernel(lambda: tr.generator_prepare(z), 20, torch)
    ms_situ = ms_plain + max(t_cp - t_gp, 0.0)
    out.append(roofline_entry(f"generator output forward with the fused interpolate (second epilogue output; in situ = plain "
                              f"product {ms_plain * 1e3:.1f} us + (critic_prepare - generator_prepare))", ms_situ,
                              2.0 * batch * G * Hg, 4.0 * (batch * Hg + Hg * G + 3 * batch * G), n_critic, pk, tf32_peak))
    dom = max(out, key=lambda r: r["ms"] * r["per_step"])
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of the same kernel + shape from the committed
    # `ncu --set full` captures (reference dims, batch 4096 only); null for shapes that were not captured
    NCU_TRAFFIC = {"critic layer-1 forward": (342.51e6 + 12.86e6, "profiles/r02_ncu_full_critic_l1_forward.txt"),
                   "critic layer-1 weight gradient": (355.25e6 + 8.50e6, "profiles/r02_ncu_full_critic_l1_wgrad.txt"),
                   "generator output forward [": (26.54e6 + 55.65e6, "profiles/r02_ncu_full_generator_output.txt"),
                   "generator output forward with the fused interpolate": (168.31e6 + 165.39e6,
                                                                           "profiles/r02_ncu_full_generator_output.txt")}
    traffic, traffic_src = None, None
    if G == 6605 and batch == 4096:
        for key, (t, src) in NCU_TRAFFIC.items():
            if dom["kernel"].startswith(key):
                traffic, traffic_src = t, src
    roof = {"bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"], "frac": dom["frac"],
            "kernel": dom["kernel"], "ms_per_launch": dom["ms"], "launches_per_step": dom["per_step"],
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": pk["src"] if dom["bound"] == "hbm" else "cuBLAS TF32 8192^3 burst measured in this run",
            "tf32_peak_measured_tflops": tf32_peak,
            "algorithmic_flops_per_launch": dom["flops"], "algorithmic_bytes_per_launch": dom["bytes"],
            "others": [{k: r[k] for k in ("kernel", "ms", "tflops", "gbs", "per_step", "bound", "frac")} for r in out]}
    return roof


def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from arshamg_scrnaseq_wgan_b200 import _lib
    _lib.build()
    ctx = dict(torch=torch, dist=dist, world=world, rank=rank, local=local, dev=f"cuda:{local}")
    if args.config == 5:
        line, keep = bench_sampling(args, ctx)
    else:
        line, keep = bench_training(args, ctx)
    if rank == 0:
        print(json.dumps(line), flush=True)
    # Teardown.  Captured graphs hold the exchange kernels: drop them first, and never let a wedged teardown keep
    # the job (and the GPUs) alive after the result line is out.
    sys.stdout.flush()
    if world > 1:
        import threading
        threading.Thread(target=lambda: (time.sleep(30), os._exit(0)), daemon=True).start()
    for tr in keep:
        tr._graph_keepalive = None
        tr._hgraph = None
    torch.cuda.synchronize()
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass
        os._exit(0)


def _barrier(ctx):
    if ctx["world"] > 1:
        ctx["dist"].barrier()
    ctx["torch"].cuda.synchronize()


def _max_over_ranks(ctx, value):
    torch = ctx["torch"]
    t = torch.tensor([value], device=ctx["dev"], dtype=torch.float64)
    if ctx["world"] > 1:
        ctx["dist"].all_reduce(t, op=ctx["dist"].ReduceOp.MAX)
    return float(t.item())


def _timed_replays(ctx, replay, steps, clocks=None):
    """K replays bracketed by barrier + synchronize; CUDA events on the launch stream; max over ranks (ms total)."""
    torch = ctx["torch"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(ctx)
    if clocks is not None and ctx["rank"] == 0:
        clocks.start()
    _barrier(ctx)
    ev0.record()
    for _ in range(steps):
        replay()
    ev1.record()
    _barrier(ctx)
    return _max_over_ranks(ctx, ev0.elapsed_time(ev1))


def _timed_host(ctx, fn, steps, warm):
    """End-to-end calls that return host values (device->host read inside): wall clock, max over ranks (seconds)."""
    for _ in range(warm):
        fn()
    _barrier(ctx)
    t0 = time.perf_counter()
    out = None
    for _ in range(steps):
        out = fn()
    _barrier(ctx)
    return _max_over_ranks(ctx, time.perf_counter() - t0), out


def bench_training(args, ctx):
    """Configs 2, 3, 4: training steps on the resident matrix, replayed as one CUDA graph."""
    torch, world, rank, local, dev = ctx["torch"], ctx["world"], ctx["rank"], ctx["local"], ctx["dev"]
    from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
    c = CONFIGS[args.config]
    dims = c["dims"]
    gen_update = args.config != 3
    if c["scaling"] == "strong":
        gbatch = args.batch or c["batch"]
        assert gbatch % world == 0
        B = gbatch // world                       # local shard of the fixed global batch
    else:
        B = args.batch or c["batch"]              # per-GPU batch, global = B * world
        gbatch = B * world
    cfg = WGANConfig(lambda_gp=10.0, n_critic=5, gp_form=args.gp_form, **dims)
    extras = args.config == 2 and not args.no_extras
    tr = WGANTrainer(cfg, max_batch=max(B, SAMPLE_CHUNK if extras else B), device=local)
    tr.init_params(seed=0)
    tr.set_data(synthetic_cells(N_CELLS, cfg.gex_size, 0, dev))
    pk = peaks()
    keep = [tr]

    # ---------------- device-resident timing: `value` ----------------
    # the whole step (device RNG, gather, critic updates, generator update, gradient exchanges) is captured once
    # into a CUDA graph; a step is one replay
    for it in range(max(args.warmup - 1, 0)):
        tr.train_iteration_resident(it, B, seed=0, generator_update=gen_update)
    replay = tr.capture_iteration(B, seed=0, start_it=max(args.warmup - 1, 0), generator_update=gen_update)
    for _ in range(2):
        replay()
    clocks = ClockSampler(local)
    ms_total = _timed_replays(ctx, replay, args.steps, clocks)
    clk = clocks.stop() if rank == 0 else None
    launches = tr.graph_launches_per_replay * args.steps
    ms_step = ms_total / args.steps
    # weak: every rank completes one batch-B step per replay -> world steps; strong: the job completes one step
    value = (world if c["scaling"] == "weak" else 1) * 1000.0 / ms_step
    last = tr.train_iteration_resident(args.warmup + args.steps + 8, B, seed=0, read=True, generator_update=gen_update)

    # ---------------- end-to-end through the public API with HOST buffers ----------------
    # (a) e2e: the framework's design point.  The training matrix is resident in HBM (uploaded once by
    #     set_data, like the reference's one-off CSV load [W:32-54]); every update's host-side inputs of the
    #     reference loop -- minibatch indices [W:188], noise [W:193], epsilon -- come from pinned host memory,
    #     the losses are read back every step.
    # (b) e2e_full_feed (config 2): additionally ship the gathered cells x [B, genes] fp32 from the host for every
    #     critic update, exactly like the reference's feed_dict [W:196,199] (PCIe-bound).
    e2e = full_feed = None
    nb = cfg.n_critic
    h2d_i = 0
    if not args.no_e2e:
        idxs = [torch.randint(N_CELLS, (B,), dtype=torch.int64).pin_memory() for _ in range(nb)]
        zs = [torch.rand((B, cfg.noise_input_size)).abs().pin_memory() for _ in range(nb + 1)]
        es = [torch.rand((B,)).pin_memory() for _ in range(nb)]
        warm = max(3, args.warmup // 2)
        dt_i, out_i = _timed_host(ctx, lambda: tr.train_iteration_host_indices_graph(idxs, zs[:nb], es, zs[nb],
                                                                                     generator_update=gen_update),
                                  args.steps, warm)
        h2d_i = sum(t_.numel() * t_.element_size() for t_ in idxs + (zs if gen_update else zs[:nb]) + es)
        mult = world if c["scaling"] == "weak" else 1
        e2e = {"value": mult * args.steps / dt_i, "unit": c["unit"], "h2d_bytes_per_step": h2d_i,
               "d2h_bytes_per_step": 32, "ms_per_step": dt_i / args.steps * 1e3,
               "api": "WGANTrainer.train_iteration_host_indices_graph(idx, z, eps, z_gen) (one CUDA graph incl. the H2D/D2H "
                      "copies): pinned host int64 minibatch indices [W:188] + float32 noise [W:193] + epsilon per critic "
                      "update, training matrix resident in HBM (set_data), device gather; losses read back every step",
               "last": out_i}
        if args.config == 2:
            xs = [torch.rand((B, cfg.gex_size)).pin_memory() for _ in range(nb)]
            dt_x, out_x = _timed_host(ctx, lambda: tr.train_iteration_host(xs, zs[:nb], es, zs[nb]), args.steps, warm)
            h2d_x = sum(t_.numel() * 4 for t_ in xs + zs + es)
            full_feed = {"value": world * args.steps / dt_x, "unit": c["unit"], "h2d_bytes_per_step": h2d_x,
                         "d2h_bytes_per_step": 32, "ms_per_step": dt_x / args.steps * 1e3,
                         "api": "WGANTrainer.train_iteration_host(x, z, eps, z_gen): x [4096,6605] fp32 fed from pinned host "
                                "memory for every critic update like feed_dict [W:196,199] (PCIe-bound)",
                         "last": out_x}
            del xs

    tf32_peak = measure_tf32_peak(torch)
    fl = flops(dims)

    # ---------------- sampler: generated cells/sec (the second half of BASELINE.json's metric) ----------------
    sampler = None
    if extras:
        SB = SAMPLE_CHUNK
        zsamp = tr.noise_prior(SB, seed=1, call_id=77)
        cells = torch.empty((SB, tr.Gl), device=dev)[:, :cfg.gex_size]
        ms_s = _max_over_ranks(ctx, time_kernel(lambda: tr.generator(zsamp, out=cells), 10, torch))
        cells_per_s = world * SB / (ms_s * 1e-3)
        sampler = {"metric": "generated_cells_per_sec", "value": cells_per_s, "unit": "cells/s", "chunk": SB,
                   "ms_per_chunk": ms_s, "note": "generator forward on one resident chunk per GPU; python bench.py --config 5 "
                                                 "times the 10 M-cell job (noise prior included) and its end-to-end form",
                   "roofline": {"bound": "tensor", "achieved": cells_per_s / world * fl["gen_fwd"] / 1e12, "peak": tf32_peak,
                                "unit": "TFLOP/s", "frac": cells_per_s / world * fl["gen_fwd"] / 1e12 / tf32_peak,
                                "per_gpu": True, "output_gbs": cells_per_s / world * cfg.gex_size * 4 / 1e9,
                                "peak_source": "cuBLAS TF32 8192^3 burst measured in this run"}}
        del cells

    # ---------------- the same step with the penalty in Gram form (reported separately) ----------------
    # layer-1 pre-activation blend + K = W1^T W1 give the same loss and gradients in real arithmetic
    # with ~30 % fewer FLOPs and no x^ / gx traffic.  `value` above stays on the explicit chain (the FLOP table's
    # denominator); this is the number a user gets with WGANConfig(gp_form="gram").
    gram = None
    if args.config == 2 and not args.no_gram and cfg.gp_form == "explicit":
        tg = WGANTrainer(WGANConfig(lambda_gp=10.0, n_critic=5, gp_form="gram", **dims), max_batch=B, device=local)
        keep.append(tg)
        tg.init_params(seed=0)
        tg._data = tr._data
        for it in range(max(args.warmup - 1, 0)):
            tg.train_iteration_resident(it, B, seed=0)
        greplay = tg.capture_iteration(B, seed=0, start_it=max(args.warmup - 1, 0))
        for _ in range(2):
            greplay()
        g_ms_step = _timed_replays(ctx, greplay, args.steps) / args.steps
        gram = {"value": world * 1000.0 / g_ms_step, "unit": c["unit"], "ms_per_step": g_ms_step,
                "gpu_launches": int(tg.graph_launches_per_replay * args.steps),
                "config": "same workload, WGANConfig(gp_form='gram'): penalty through K = W1^T W1 (SURVEY 8d shortcuts i+ii)",
                "losses": tg.train_iteration_resident(args.warmup + args.steps + 8, B, seed=0, read=True)}
        if not args.no_e2e:
            dt_g, out_g = _timed_host(ctx, lambda: tg.train_iteration_host_indices_graph(idxs, zs[:nb], es, zs[nb]),
                                      args.steps, max(3, args.warmup // 2))
            gram["e2e"] = {"value": world * args.steps / dt_g, "unit": c["unit"], "ms_per_step": dt_g / args.steps * 1e3,
                           "h2d_bytes_per_step": h2d_i, "d2h_bytes_per_step": 32, "last": out_g}
        tg._graph_keepalive = None
        tg._hgraph = None
        torch.cuda.synchronize()

    # ---------------- the reference's own configuration: batch 32, no penalty, 1 generator + 1 critic update ----------
    literal = small = None
    if extras and world == 1:
        tl = WGANTrainer(WGANConfig.reference_literal(), max_batch=32, device=local)
        keep.append(tl)
        tl.init_params(seed=0)
        tl._data, tl._data_max_value = tr._data, tr._data_max_value
        lreplay = tl.capture_iteration(32, seed=0)
        l_ms = _timed_replays(ctx, lreplay, 200) / 200
        literal = {"what": "scripts/WGAN-GP_minimal.py as written: batch 32 [W:11], no penalty, generator update then critic "
                           "update per iteration [W:196-199], one CUDA-graph replay per iteration",
                   "gpu_iters_per_sec": 1000.0 / l_ms, "gpu_ms_per_iter": l_ms,
                   "gpu_launches_per_iter": int(tl.graph_launches_per_replay)}
        tl._graph_keepalive = None
        # ... and WGAN-GP at the reference's batch size (lambda 10, n_critic 5, batch 32)
        ts = WGANTrainer(WGANConfig(lambda_gp=10.0, n_critic=5, **dims), max_batch=32, device=local)
        keep.append(ts)
        ts.init_params(seed=0)
        ts._data, ts._data_max_value = tr._data, tr._data_max_value
        sreplay = ts.capture_iteration(32, seed=0)
        s_ms = _timed_replays(ctx, sreplay, 200) / 200
        small = {"what": "WGAN-GP iteration (5 critic+penalty updates + 1 generator update) at the reference's batch 32 [W:11]",
                 "gpu_iters_per_sec": 1000.0 / s_ms, "gpu_ms_per_iter": s_ms,
                 "gpu_launches_per_iter": int(ts.graph_launches_per_replay)}
        ts._graph_keepalive = None

    roof = cpu = None
    if rank == 0:
        roof = dominant_kernel_roofline(tr, B, torch, pk, tf32_peak, cfg.n_critic, gen_update)
        if world == 1 and not args.no_cpu_baseline:
            sps, ms, cores, sample = time_cpu(args.config, gbatch, 5, 2, budget_s=40.0)
            cpu = {"value": cpu_value(args.config, sps, gbatch), "unit": c["unit"], "cores": cores, "kind": "port",
                   "sample": sample}
            if sampler is not None:
                sps5, _, cores5, sample5 = time_cpu(5, 4096, 5, 2, budget_s=10.0)
                sampler["cpu_baseline"] = {"value": sps5 * 4096, "unit": "cells/s", "cores": cores5, "kind": "port",
                                           "sample": sample5}
            if literal is not None:
                spsl, msl, coresl, samplel = time_cpu(2, 32, 50, 5, budget_s=30.0, literal=True)
                literal.update({"cpu_iters_per_sec": spsl, "cpu_ms_per_iter": msl, "cpu_cores": coresl, "cpu_sample": samplel})
                spss, mss, coress, samples = time_cpu(2, 32, 20, 3, budget_s=30.0)
                small.update({"cpu_iters_per_sec": spss, "cpu_ms_per_iter": mss, "cpu_cores": coress, "cpu_sample": samples})
    line = {"metric": c["metric"], "value": value, "unit": c["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": c["scaling"],
            "vs_baseline": None, "dtype": "tf32 (fp32 storage, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": c["workload"], "baseline_config": args.config,
                       "per_gpu_batch": B, "global_batch": gbatch, "parallelism": f"dp{world}",
                       "penalty_form": cfg.gp_form,
                       "execution": "training matrix resident in HBM, device Philox RNG + gather, the step replayed as one CUDA graph",
                       "l2": "per-step working set > 126 MB L2 (inputs larger than L2)",
                       "step_tflops": (gbatch * c["flop_per_sample_step"] / (ms_step * 1e-3) / 1e12)},
            "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu,
            "losses": last}
    if full_feed is not None:
        line["e2e_full_feed"] = full_feed
    if sampler is not None:
        line["sampler"] = sampler
    if gram is not None:
        line["gram_form"] = gram
    if literal is not None:
        line["reference_literal"] = literal
        line["wgan_gp_batch32"] = small
    return line, keep


def bench_sampling(args, ctx):
    """Config 5: 10 M generated cells sharded over the ranks (a step = the whole job) + the latent-vector fit."""
    torch, world, rank, local, dev = ctx["torch"], ctx["world"], ctx["rank"], ctx["local"], ctx["dev"]
    from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
    c = CONFIGS[5]
    cfg = WGANConfig(**c["dims"])
    chunk = args.batch or SAMPLE_CHUNK
    tr = WGANTrainer(cfg, max_batch=chunk, device=local)
    tr.init_params(seed=0)
    pk = peaks()
    n_local = (SAMPLE_CELLS + world - 1) // world
    nchunks = [0]

    def sink(row0, cells):                     # the cells stay on the device (a consumer kernel would read them here)
        nchunks[0] += 1

    def job():
        tr.sample_cells(n_local, chunk=chunk, seed=7, sink=sink, first_row=rank * n_local)

    for _ in range(max(1, min(args.warmup, 2))):
        job()
    l0 = tr.launch_count
    clocks = ClockSampler(local)
    steps = args.steps
    ms_total = _timed_replays(ctx, job, steps, clocks)
    clk = clocks.stop() if rank == 0 else None
    launches = tr.launch_count - l0
    ms_step = ms_total / steps
    value = SAMPLE_CELLS / (ms_step * 1e-3)

    # e2e: the reference's sampling contract is host noise in, host cells out [W:217-221].  Cells leave through a sink
    # that copies every chunk into pinned host memory (double-buffered on a copy stream); the latent noise of every
    # chunk comes from pinned host memory.  Bounded: E2E_CELLS per step per GPU (PCIe-bound: 26.4 KB per cell out).
    e2e = None
    if not args.no_e2e:
        E2E_CELLS = 8 * chunk
        zs_h = [torch.rand((chunk, cfg.noise_input_size)).abs().pin_memory() for _ in range(2)]
        out_h = [torch.empty((chunk, cfg.gex_size)).pin_memory() for _ in range(2)]
        z_d = [torch.empty((chunk, cfg.noise_input_size), device=dev) for _ in range(2)]
        c_d = [torch.empty((chunk, tr.Gl), device=dev) for _ in range(2)]
        copy = torch.cuda.Stream(device=dev)
        done = [torch.cuda.Event() for _ in range(2)]

        def e2e_job():
            main = torch.cuda.current_stream()
            for i in range(E2E_CELLS // chunk):
                s = i % 2
                main.wait_event(done[s])                       # the D2H copy that last used this slot is finished
                z_d[s].copy_(zs_h[s], non_blocking=True)
                tr.generator(z_d[s], out=c_d[s][:, :cfg.gex_size])
                copy.wait_stream(main)
                with torch.cuda.stream(copy):
                    out_h[s].copy_(c_d[s][:, :cfg.gex_size], non_blocking=True)
                    done[s].record(copy)
            copy.synchronize()
            return float(out_h[0][0, 0])

        esteps = max(2, min(steps, 5))
        dt, _ = _timed_host(ctx, e2e_job, esteps, 1)
        e2e = {"value": world * E2E_CELLS * esteps / dt, "unit": c["unit"],
               "h2d_bytes_per_step": E2E_CELLS * cfg.noise_input_size * 4, "d2h_bytes_per_step": E2E_CELLS * cfg.gex_size * 4,
               "ms_per_step": dt / esteps * 1e3, "steps": esteps,
               "api": f"WGANTrainer.generator(z, out) per chunk of {chunk} cells: latent noise from pinned host memory, cells "
                      f"copied to pinned host memory on a second stream; bounded sample of {E2E_CELLS} cells per GPU and step "
                      "(26.4 KB per cell device->host: PCIe-bound)"}

    tf32_peak = measure_tf32_peak(torch)
    fl = flops(c["dims"])
    per_gpu = value / world
    roof = {"bound": "tensor", "achieved": per_gpu * fl["gen_fwd"] / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
            "frac": per_gpu * fl["gen_fwd"] / 1e12 / tf32_peak,
            "kernel": "the whole sampling job per GPU (noise prior + 3 generator GEMMs per chunk; 93 % of the FLOPs are the "
                      "output product [chunk x 600] x [600 x 6605], TF32 tcgen05 cta_group::2)",
            "ms_per_launch": ms_step / max(1, (n_local + chunk - 1) // chunk), "traffic": None,
            "peak_source": "cuBLAS TF32 8192^3 burst measured in this run", "tf32_peak_measured_tflops": tf32_peak,
            "output_gbs_per_gpu": per_gpu * cfg.gex_size * 4 / 1e9,
            "algorithmic_flops_per_launch": chunk * fl["gen_fwd"],
            "algorithmic_bytes_per_launch": 4.0 * chunk * (cfg.noise_input_size + cfg.gex_size)}

    # latent-vector TF-Adam fit (scripts/Adam_prediction.py's optimizer on z): cells sharded, no exchange
    LF = chunk
    target = tr.sample_cells(LF, seed=3, first_row=rank * LF)
    z, loss = tr.latent_fit(target, steps=3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(ctx)
    e0.record()
    z, loss = tr.latent_fit(target, steps=20, z0=z)
    e1.record()
    _barrier(ctx)
    lf_ms = _max_over_ranks(ctx, e0.elapsed_time(e1)) / 20
    latent = {"what": f"one TF-Adam step on the latent vectors of {LF} target cells per GPU (generator frozen)",
              "ms_per_adam_step": lf_ms, "cell_steps_per_sec": world * LF / (lf_ms * 1e-3), "loss_mean": float(loss.mean())}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, ms, cores, sample = time_cpu(5, 4096, 5, 2, budget_s=20.0)
        cpu = {"value": sps * 4096, "unit": c["unit"], "cores": cores, "kind": "port", "sample": sample}
    line = {"metric": c["metric"], "value": value, "unit": c["unit"], "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "tf32 (fp32 storage, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": c["workload"], "baseline_config": 5, "cells_per_step": SAMPLE_CELLS,
                       "cells_per_gpu": n_local, "chunk": chunk, "parallelism": f"sharded x{world}, no communication",
                       "l2": "each chunk writes 866 MB > 126 MB L2"},
            "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu,
            "latent_fit": latent}
    return line, [tr]


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

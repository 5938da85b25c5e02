"""Host-side mirror of the reference script's interface for the hot path.

Names follow scripts/WGAN-GP_minimal.py: the module constants [W:9-24] are ``WGANConfig``
fields, ``generator`` / ``discriminator`` / ``noise_prior`` keep their names and argument
meaning [W:61,91,103], variables keep their TF names [W:63-122], and the optimizer keeps the
``Adam_Prediction_Optimizer`` signature [A:38-39].  All arithmetic happens in the sm_100a CUDA
library behind include/wgan_b200.h (loaded through ctypes); PyTorch is used only for device
memory, streams and ``torch.distributed``.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
import dataclasses
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib

PARAM_NAMES = (
    "generator/gen_dense1/kernel", "generator/gen_dense1/bias",
    "generator/gen_dense2/kernel", "generator/gen_dense2/bias",
    "generator/gen_output/kernel", "generator/gen_output/bias",
    "discriminator/disc_dense1/kernel", "discriminator/disc_dense1/bias",
    "discriminator/disc_dense2/kernel", "discriminator/disc_dense2/bias",
    "discriminator/disc_output/kernel", "discriminator/disc_output/bias",
)
GEN_NAMES, CRITIC_NAMES = PARAM_NAMES[:6], PARAM_NAMES[6:]
NET_GENERATOR, NET_CRITIC = 0, 1
KIND_PARAM, KIND_GRAD, KIND_M, KIND_V = 0, 1, 2, 3
KINDS = {"param": 0, "grad": 1, "m": 2, "v": 3}


@dataclasses.dataclass
class Adam_Prediction_Optimizer:
    """Hyper-parameter record with the reference constructor's signature [A:38-39].

    ``prediction`` is accepted for API parity; all gradients of this model are dense, so the
    reference's prediction branch (sparse path only, [A:189-194]) never runs."""
    learning_rate: float = 0.001
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-8
    prediction: bool = False
    use_locking: bool = False
    name: str = "Adam"


@dataclasses.dataclass
class WGANConfig:
    n_train_steps: int = 10000          # [W:9]
    batch_size: int = 32                # [W:11]
    noise_input_size: int = 100         # [W:13]
    inflate_to_size: int = 600          # [W:14]
    gex_size: int = 6605                # [W:15]
    disc_internal_size: int = 200       # [W:17]
    num_cells_train: int = 1500         # [W:18]
    num_cells_generate: int = 500       # [W:20]
    learn_rate: float = 1e-5            # [W:22]
    beta1: float = 0.9                  # [W:154-155]
    beta2: float = 0.999
    epsilon: float = 1e-8               # [A:38]
    alpha: float = 0.2                  # tf.nn.leaky_relu default
    lambda_gp: float = 10.0             # build-defined; 0 => reference-literal
    gp_form: str = "explicit"           # "explicit": x^ materialised, gx = g1 W1^T;
                                        # "gram": same penalty through K = W1^T W1 (SURVEY 8d shortcuts)
    n_critic: int = 5                   # build-defined; 1 => reference-literal
    step_order: str = "critic_first"    # "gen_first" => reference-literal [W:196,199]

    @classmethod
    def reference_literal(cls, **kw) -> "WGANConfig":
        """Exactly what the reference script computes: no penalty, 1:1 updates, generator first."""
        return cls(lambda_gp=0.0, n_critic=1, step_order="gen_first", **kw)


class _DevPtr:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned memory."""

    def __init__(self, ptr: int, n: int, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False),
                                         "version": 2}


class _OnDevice:
    """The handle-less input-pipeline entry points launch on the CURRENT CUDA device; this proxy makes the
    trainer's device current around each call (two trainers on different GPUs in one process)."""

    def __init__(self, lib, device):
        self._lib, self._device = lib, device

    def __getattr__(self, name):
        fn = getattr(self._lib, name)

        def call(*a):
            with torch.cuda.device(self._device):
                return fn(*a)
        return call


def _stream(device=None) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class WGANTrainer:
    """One handle per GPU / per process.  Data-parallel when torch.distributed is initialised (or when
    ``dp_rank`` / ``dp_world`` are given and ``dp_connect`` is called by hand): every rank passes its batch shard,
    gradients are SUM-all-reduced -- one exchange per optimizer step over the flat gradient arena, done by the
    library's own kernel over NVLink peer memory (``wgan_allreduce_grads``) -- and Adam runs replicated.
    ``torch.distributed`` only carries the 192-byte rendezvous blobs.  ``dp_backend="nccl"`` keeps the
    round-1 ``torch.distributed.all_reduce`` exchange for A/B measurements."""

    def __init__(self, cfg: WGANConfig, max_batch: Optional[int] = None, device: Optional[int] = None,
                 gemm_mode: int = 0, optimizer: Optional[Adam_Prediction_Optimizer] = None,
                 process_group=None, distributed: Optional[bool] = None, dp_rank: Optional[int] = None,
                 dp_world: Optional[int] = None, dp_backend: Optional[str] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("WGANTrainer needs a CUDA device (B200); there is no CPU fallback")
        self.cfg = cfg
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        torch.cuda.set_device(self.device)
        self.max_batch = int(max_batch if max_batch is not None else cfg.batch_size)
        if optimizer is not None:
            cfg = dataclasses.replace(cfg, learn_rate=optimizer.learning_rate, beta1=optimizer.beta1,
                                      beta2=optimizer.beta2, epsilon=optimizer.epsilon)
            self.cfg = cfg
        if cfg.gp_form not in ("explicit", "gram"):
            raise ValueError(f"gp_form must be 'explicit' or 'gram', got {cfg.gp_form!r}")
        # data parallel plumbing: rank / world come from torch.distributed unless given explicitly
        import torch.distributed as dist
        self.dist = dist
        use_dist = (dist.is_available() and dist.is_initialized()) if distributed is None else distributed
        self.group = process_group
        if dp_world is not None:
            self.world, self.rank = int(dp_world), int(dp_rank or 0)
            use_dist = False
        else:
            self.world = dist.get_world_size(process_group) if use_dist else 1
            self.rank = dist.get_rank(process_group) if use_dist else 0
        self.dp_backend = dp_backend or os.environ.get("WGAN_B200_DP_BACKEND", "engine")
        if self.dp_backend not in ("engine", "nccl"):
            raise ValueError(f"dp_backend must be 'engine' or 'nccl', got {self.dp_backend!r}")
        c = _lib.WganConfig(cfg.noise_input_size, cfg.inflate_to_size, cfg.gex_size, cfg.disc_internal_size,
                            self.max_batch, self.device, gemm_mode, 1 if cfg.gp_form == "gram" else 0,
                            cfg.learn_rate, cfg.beta1, cfg.beta2,
                            cfg.epsilon, cfg.alpha, cfg.lambda_gp, self.rank, self.world)
        self._h = C.c_void_p()
        rc = self.lib.wgan_create(C.byref(c), C.byref(self._h))
        if rc != 0:
            raise _lib.WganError("wgan_create failed: " + self.lib.wgan_last_error(None).decode())
        self.info = {}
        for i in range(self.lib.wgan_tensor_count()):
            ti = _lib.WganTensorInfo()
            self._ck(self.lib.wgan_get_tensor_info(self._h, i, C.byref(ti)), "get_tensor_info")
            self.info[ti.name.decode()] = (i, ti.net, ti.rows, ti.cols, ti.ld, ti.offset)
        self.Gl = self.lib.wgan_padded_ld(cfg.gex_size)
        self._grad_arena = {n: self.arena(n, KIND_GRAD) for n in (NET_GENERATOR, NET_CRITIC)}
        self._pl = _OnDevice(self.lib, self.device)
        self._data = None
        self._data_max_value = 1.0
        self._calls = 0
        self._dp_connected = self.world == 1
        self._comm_stream = None
        if self.world > 1 and use_dist and self.dp_backend == "engine":
            self._dp_connect_distributed()

    # ------------------------------------------------------------------ plumbing
    def _st(self) -> C.c_void_p:
        """torch's current stream ON THIS TRAINER'S GPU (not the process-wide current device)."""
        return _stream(self.device)

    def _ck(self, rc, what):
        _lib.check(self.lib, self._h, rc, what)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.wgan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def arena(self, net: int, kind: int) -> torch.Tensor:
        """Flat fp32 view of one arena (library-owned memory)."""
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.wgan_get_arena(self._h, net, kind, C.byref(p), C.byref(n)), "get_arena")
        return torch.as_tensor(_DevPtr(p.value, n.value), device=f"cuda:{self.device}")

    def buffer(self, name: str) -> torch.Tensor:
        p, r, c, ld = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.wgan_get_buffer(self._h, name.encode(), C.byref(p), C.byref(r), C.byref(c), C.byref(ld)),
                 "get_buffer")
        flat = torch.as_tensor(_DevPtr(p.value, r.value * ld.value), device=f"cuda:{self.device}")
        return flat.view(r.value, ld.value)[:, :c.value]

    @property
    def launch_count(self) -> int:
        return int(self.lib.wgan_launch_count(self._h))

    # ------------------------------------------------------------------ state I/O
    def get_tensor(self, name: str, kind: str = "param") -> np.ndarray:
        i, _, rows, cols, _, _ = self.info[name]
        out = np.empty((rows, cols), np.float32)
        self._ck(self.lib.wgan_get_tensor(self._h, KINDS[kind], i, out.ctypes.data_as(C.c_void_p)), "get_tensor")
        return out[0] if name.endswith("bias") else out

    def set_tensor(self, name: str, value, kind: str = "param"):
        i, _, rows, cols, _, _ = self.info[name]
        v = np.ascontiguousarray(np.asarray(value, np.float32).reshape(rows, cols))
        self._ck(self.lib.wgan_set_tensor(self._h, KINDS[kind], i, v.ctypes.data_as(C.c_void_p)), "set_tensor")

    def set_params(self, params: Dict[str, np.ndarray]):
        for k, v in params.items():
            self.set_tensor(k, v)

    def get_params(self) -> Dict[str, np.ndarray]:
        return {k: self.get_tensor(k) for k in PARAM_NAMES}

    def get_grads(self, names=PARAM_NAMES) -> Dict[str, np.ndarray]:
        return {k: self.get_tensor(k, "grad") for k in names}

    def adam_powers(self, net: int):
        a, b = C.c_float(), C.c_float()
        self._ck(self.lib.wgan_get_adam_powers(self._h, net, C.byref(a), C.byref(b)), "get_adam_powers")
        return a.value, b.value

    def set_adam_powers(self, net: int, b1p: float, b2p: float):
        self._ck(self.lib.wgan_set_adam_powers(self._h, net, b1p, b2p), "set_adam_powers")

    def init_params(self, seed: int = 0):
        """tf.layers.dense defaults: Glorot-uniform kernels, zero biases [W:65-80,107-122]."""
        rng = np.random.default_rng(seed)
        for name in PARAM_NAMES:
            _, _, rows, cols, _, _ = self.info[name]
            if name.endswith("kernel"):
                lim = np.sqrt(6.0 / (rows + cols))
                self.set_tensor(name, rng.uniform(-lim, lim, size=(rows, cols)))
            else:
                self.set_tensor(name, np.zeros((rows, cols)))

    def state_dict(self) -> dict:
        """Everything tf.train.Saver would save [W:171]: variables, Adam slots, beta powers."""
        sd = {"config": dataclasses.asdict(self.cfg)}
        for kind in ("param", "m", "v"):
            sd[kind] = {k: self.get_tensor(k, kind) for k in PARAM_NAMES}
        sd["beta_powers"] = {"generator": self.adam_powers(NET_GENERATOR), "critic": self.adam_powers(NET_CRITIC)}
        return sd

    def load_state_dict(self, sd: dict):
        for kind in ("param", "m", "v"):
            for k, v in sd[kind].items():
                self.set_tensor(k, v, kind)
        self.set_adam_powers(NET_GENERATOR, *sd["beta_powers"]["generator"])
        self.set_adam_powers(NET_CRITIC, *sd["beta_powers"]["critic"])

    # ------------------------------------------------------------------ steps
    def _mat(self, t: torch.Tensor, cols: int):
        assert t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape[1] == cols, \
            f"expected a cuda float32 [rows, {cols}] tensor, got {tuple(t.shape)} {t.dtype} {t.device}"
        assert t.stride(1) == 1
        return C.c_void_p(t.data_ptr()), int(t.stride(0)), int(t.shape[0])

    # ------------------------------------------------------------------ data parallelism
    def dp_export(self) -> bytes:
        """This rank's rendezvous blob (wgan_dp_export): hand the blobs of all ranks, in rank order, to dp_connect."""
        buf = C.create_string_buffer(_lib.DP_BLOB_BYTES)
        self._ck(self.lib.wgan_dp_export(self._h, buf), "dp_export")
        return buf.raw

    def dp_connect(self, blobs):
        """Map the gradient windows of all ranks (wgan_dp_connect).  `blobs`: one bytes object per rank."""
        assert len(blobs) == self.world
        joined = b"".join(blobs)
        self._ck(self.lib.wgan_dp_connect(self._h, C.c_char_p(joined)), "dp_connect")
        self._dp_connected = True

    def _dp_connect_distributed(self):
        mine = torch.frombuffer(bytearray(self.dp_export()), dtype=torch.uint8)
        backend = self.dist.get_backend(self.group)
        if backend == "nccl":
            mine = mine.to(f"cuda:{self.device}")
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(parts, mine, group=self.group)
        self.dp_connect([bytes(t.cpu().numpy().tobytes()) for t in parts])
        self.dist.barrier(group=self.group)      # every rank has mapped every window before the first exchange

    def _allreduce(self, net: int, overlap=None):
        """One SUM all-reduce of the flat gradient arena (+ loss scalars) per optimizer step
        (wgan_allreduce_grads: the library's own kernel over NVLink peer memory).  `overlap` is host code that
        enqueues work independent of the reduced gradients (the next update's input pipeline and generator
        forward); with overlap the exchange runs on a second stream next to it."""
        if self.world == 1:
            if overlap is not None:
                overlap()
            return
        if self.dp_backend == "nccl":
            work = self.dist.all_reduce(self._grad_arena[net], op=self.dist.ReduceOp.SUM, group=self.group,
                                        async_op=True)
            if overlap is not None:
                overlap()
            work.wait()
            return
        assert self._dp_connected, "data-parallel trainer: call dp_connect(blobs) before the first step"
        if overlap is None or os.environ.get("WGAN_B200_DP_SERIAL"):
            self._ck(self.lib.wgan_allreduce_grads(self._h, net, self._st()), "allreduce_grads")
            if overlap is not None:
                overlap()
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=f"cuda:{self.device}", priority=-1)
        main = torch.cuda.current_stream(self.device)
        self._comm_stream.wait_stream(main)
        self._ck(self.lib.wgan_allreduce_grads(self._h, net, C.c_void_p(self._comm_stream.cuda_stream)),
                 "allreduce_grads")
        overlap()
        main.wait_stream(self._comm_stream)

    def critic_prepare(self, x, z, eps=None) -> int:
        """Run the critic-parameter-independent part of the next critic_step(x, z, eps) now; the returned token is
        honoured by that critic_step as long as nothing else used the activation workspace in between."""
        xp, ldx, rows = self._mat(x, self.cfg.gex_size)
        zp, ldz, _ = self._mat(z, self.cfg.noise_input_size)
        ep = C.c_void_p(eps.data_ptr()) if eps is not None else C.c_void_p()
        tok = C.c_uint64(0)
        self._ck(self.lib.wgan_critic_prepare(self._h, xp, ldx, zp, ldz, ep, rows, self._st(), C.byref(tok)),
                 "critic_prepare")
        return tok.value

    def scalars(self, net: int) -> np.ndarray:
        out = np.empty(4, np.float32)
        self._ck(self.lib.wgan_get_scalars(self._h, net, out.ctypes.data_as(C.c_void_p), self._st()), "get_scalars")
        return out

    def generator_prepare(self, z) -> int:
        """Run the critic-independent part (z staging, G(z)) of the next generator_step(z) now."""
        zp, ldz, rows = self._mat(z, self.cfg.noise_input_size)
        tok = C.c_uint64(0)
        self._ck(self.lib.wgan_generator_prepare(self._h, zp, ldz, rows, self._st(), C.byref(tok)), "generator_prepare")
        return tok.value

    def critic_step(self, x, z, eps=None, apply: bool = True, read: bool = True, global_batch: Optional[int] = None,
                    overlap=None, token: Optional[int] = None):
        """One critic update [W:199]: gradients of obj_d (+GP), all-reduce, Adam.
        x [rows, gex], z [rows, noise], eps [rows] are this rank's shard.  `token`: what critic_prepare returned
        for exactly these inputs, unchanged since (the library ignores a stale token and redoes the work)."""
        xp, ldx, rows = self._mat(x, self.cfg.gex_size)
        zp, ldz, rz = self._mat(z, self.cfg.noise_input_size)
        assert rz == rows
        ep = C.c_void_p(eps.data_ptr()) if eps is not None else C.c_void_p()
        if eps is not None:
            assert eps.is_cuda and eps.dtype == torch.float32 and eps.numel() == rows and eps.is_contiguous()
        gb = rows * self.world if global_batch is None else int(global_batch)
        self._ck(self.lib.wgan_critic_grads(self._h, xp, ldx, zp, ldz, ep, rows, gb, int(token or 0), self._st()),
                 "critic_grads")
        self._allreduce(NET_CRITIC, overlap)
        if apply:
            self._ck(self.lib.wgan_critic_apply(self._h, self._st()), "critic_apply")
        if not read:
            return None
        s = self.scalars(NET_CRITIC)
        wass = float(s[0] - s[1])                                       # obj_d [W:137]
        return {"wasserstein": wass, "gp": float(s[2]), "critic_loss": wass + float(s[2])}

    def generator_step(self, z, apply: bool = True, read: bool = True, global_batch: Optional[int] = None,
                       token: Optional[int] = None):
        """One generator update [W:196]: gradients of obj_g through the critic, all-reduce, Adam."""
        zp, ldz, rows = self._mat(z, self.cfg.noise_input_size)
        gb = rows * self.world if global_batch is None else int(global_batch)
        if self.world > 1 and self.dp_backend == "engine" and not os.environ.get("WGAN_B200_DP_SERIAL"):
            # gradients + exchange in one call: the output layer's 16 MB go out while the backward pass continues
            assert self._dp_connected, "data-parallel trainer: call dp_connect(blobs) before the first step"
            self._ck(self.lib.wgan_generator_grads_exchange(self._h, zp, ldz, rows, gb, int(token or 0), self._st()),
                     "generator_grads_exchange")
        else:
            self._ck(self.lib.wgan_generator_grads(self._h, zp, ldz, rows, gb, int(token or 0), self._st()),
                     "generator_grads")
            self._allreduce(NET_GENERATOR)
        if apply:
            self._ck(self.lib.wgan_generator_apply(self._h, self._st()), "generator_apply")
        if not read:
            return None
        return {"gen_loss": -float(self.scalars(NET_GENERATOR)[0])}     # obj_g [W:138]

    def train_iteration(self, x, z, eps=None, read: bool = True):
        """Reference-literal iteration [W:188-199]: generator update, then critic update, on the
        same batch and noise (the critic step sees the updated generator)."""
        g = self.generator_step(z, read=read)
        c = self.critic_step(x, z, eps, read=read)
        return {**g, **c} if read else None

    # ------------------------------------------------------------------ model functions
    def generator(self, z, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """G(z) [W:61-83]: [rows, noise] -> [rows, gex]; batched (any rows <= max_batch)."""
        zp, ldz, rows = self._mat(z, self.cfg.noise_input_size)
        if out is None:
            out = torch.empty((rows, self.Gl), device=z.device, dtype=torch.float32)[:, :self.cfg.gex_size]
        op, ldo, ro = self._mat(out, self.cfg.gex_size)
        assert ro == rows
        self._ck(self.lib.wgan_sample(self._h, zp, ldz, op, ldo, rows, self._st()), "sample")
        return out

    def discriminator(self, x) -> torch.Tensor:
        """D(x) [W:103-124]: [rows, gex] -> [rows, 1] (linear output, no sigmoid)."""
        xp, ldx, rows = self._mat(x, self.cfg.gex_size)
        out = torch.empty((rows, 1), device=x.device, dtype=torch.float32)
        self._ck(self.lib.wgan_critic_forward(self._h, xp, ldx, C.c_void_p(out.data_ptr()), rows, self._st()),
                 "critic_forward")
        return out

    def noise_prior(self, batch_size: int, dim: Optional[int] = None, seed: int = 0, call_id: Optional[int] = None,
                    row0: Optional[int] = None, data_max_value: Optional[float] = None, counter=None) -> torch.Tensor:
        """abs(Normal(0, data_max/10) + Poisson(1)) [W:91-94], drawn on the device (Philox).  data_max_value defaults
        to the maximum of the resident training matrix (set_data), 1.0 before any data is set."""
        data_max_value = self._data_max_value if data_max_value is None else data_max_value
        dim = self.cfg.noise_input_size if dim is None else dim
        if call_id is None:
            call_id = self._calls
            self._calls += 1
        if row0 is None:
            row0 = self.rank * batch_size
        out = torch.empty((batch_size, dim), device=f"cuda:{self.device}", dtype=torch.float32)
        rc = self._pl.wgan_rng_noise_prior(C.c_void_p(out.data_ptr()), dim, row0, batch_size, dim,
                                           data_max_value / 10.0, seed, call_id, self._cptr(counter), self._st())
        if rc != 0:
            raise _lib.WganError("rng_noise_prior: " + self.lib.wgan_last_error(None).decode())
        return out

    @staticmethod
    def _cptr(counter):
        return C.c_void_p(counter.data_ptr()) if counter is not None else C.c_void_p()

    def uniform(self, rows: int, seed: int, call_id: int, row0: Optional[int] = None, counter=None) -> torch.Tensor:
        out = torch.empty((rows,), device=f"cuda:{self.device}", dtype=torch.float32)
        row0 = self.rank * rows if row0 is None else row0
        rc = self._pl.wgan_rng_uniform(C.c_void_p(out.data_ptr()), row0, rows, seed, call_id, self._cptr(counter),
                                       self._st())
        if rc != 0:
            raise _lib.WganError("rng_uniform: " + self.lib.wgan_last_error(None).decode())
        return out

    def indices(self, rows: int, n_cells: int, seed: int, call_id: int, row0: Optional[int] = None,
                counter=None) -> torch.Tensor:
        out = torch.empty((rows,), device=f"cuda:{self.device}", dtype=torch.int64)
        row0 = self.rank * rows if row0 is None else row0
        rc = self._pl.wgan_rng_indices(C.c_void_p(out.data_ptr()), row0, rows, n_cells, seed, call_id,
                                       self._cptr(counter), self._st())
        if rc != 0:
            raise _lib.WganError("rng_indices: " + self.lib.wgan_last_error(None).decode())
        return out

    # ------------------------------------------------------------------ resident-data training
    def set_data(self, cells_by_genes: torch.Tensor, data_max_value: Optional[float] = None):
        """Keep the (already min-max scaled) training matrix resident, cells-major, padded rows
        (the reference keeps it genes-major on the host and gathers columns every step [W:188-191]).
        `data_max_value` = np.amax(training matrix) [W:89] (what prepare_training_matrix returns; computed here when
        omitted): the noise prior of every resident / graph / sampling path uses sigma = data_max_value / 10 [W:93]."""
        n, g = cells_by_genes.shape
        assert g == self.cfg.gex_size
        buf = torch.zeros((n, self.Gl), device=f"cuda:{self.device}", dtype=torch.float32)
        buf[:, :g] = cells_by_genes.to(buf.device, torch.float32)
        self._data = buf
        self._data_max_value = float(buf.max().item()) if data_max_value is None else float(data_max_value)

    def real_slot(self, rows: int) -> torch.Tensor:
        """The rows of the library's stacked activation buffer ([fake | interpolate | real], or
        [fake | real] without a penalty and in the Gram form) where the critic step expects the real
        cells; gathering straight into it lets layer 1 and its weight gradient run as single GEMMs
        over the stacked rows (no staging copy)."""
        xcat = self.buffer("xcat")
        r0 = (2 if (self.cfg.lambda_gp != 0 and self.cfg.gp_form == "explicit") else 1) * rows
        return xcat[r0:r0 + rows]

    def gather(self, idx: torch.Tensor) -> torch.Tensor:
        rows = idx.numel()
        out = self.real_slot(rows)
        rc = self._pl.wgan_gather_rows(C.c_void_p(self._data.data_ptr()), self.Gl, C.c_void_p(idx.data_ptr()),
                                       C.c_void_p(out.data_ptr()), self.Gl, rows, self.cfg.gex_size, self._st())
        if rc != 0:
            raise _lib.WganError("gather_rows: " + self.lib.wgan_last_error(None).decode())
        return out

    def draw_minibatch(self, batch: int, seed: int, call_id: int, counter=None, data_max_value: Optional[float] = None):
        """Indices [W:188] + gather [W:189-191] + noise [W:193] + epsilon of one critic update in a single
        launch; the same values as indices() / gather() / noise_prior() / uniform() with this call id."""
        dev = f"cuda:{self.device}"
        data_max_value = self._data_max_value if data_max_value is None else data_max_value
        x = self.real_slot(batch)
        z = torch.empty((batch, self.cfg.noise_input_size), device=dev, dtype=torch.float32)
        eps = torch.empty((batch,), device=dev, dtype=torch.float32) if self.cfg.lambda_gp != 0 else None
        rc = self._pl.wgan_draw_minibatch(
            C.c_void_p(self._data.data_ptr()), self.Gl, self._data.shape[0], C.c_void_p(x.data_ptr()), self.Gl,
            self.cfg.gex_size, C.c_void_p(z.data_ptr()), z.shape[1], z.shape[1], data_max_value / 10.0,
            C.c_void_p(eps.data_ptr()) if eps is not None else C.c_void_p(), C.c_void_p(), self.rank * batch, batch,
            seed, call_id, self._cptr(counter), self._st())
        if rc != 0:
            raise _lib.WganError("draw_minibatch: " + self.lib.wgan_last_error(None).decode())
        return x, z, eps

    def train_iteration_resident(self, it: int, batch: int, seed: int = 0, read: bool = False, counter=None,
                                 generator_update: bool = True):
        """One WGAN-GP iteration with everything drawn on the device: n_critic critic updates
        (fresh minibatch, noise and epsilon each) and one generator update (canonical order),
        or the reference-literal order when cfg.step_order == 'gen_first'.  RNG call ids are
        it * (n_critic + 1) + k; with `counter` (a device uint32) the base comes from the device.
        generator_update=False runs the n_critic loop alone (BASELINE config 3, the gradient-penalty stress)."""
        cfg = self.cfg
        base = 0 if counter is not None else it * (cfg.n_critic + 1)
        out = {}
        if cfg.step_order == "gen_first":
            x, z, eps = self.draw_minibatch(batch, seed, base, counter=counter)
            return self.train_iteration(x, z, eps, read=read)

        def draw(k):
            if os.environ.get("WGAN_B200_SPLIT_DRAW"):          # experiment: the four separate launches
                cid = base + k
                z = self.noise_prior(batch, seed=seed, call_id=cid, counter=counter,
                                     data_max_value=self._data_max_value)
                x = self.gather(self.indices(batch, self._data.shape[0], seed, cid, counter=counter))
                eps = self.uniform(batch, seed, cid, counter=counter) if cfg.lambda_gp != 0 else None
                return x, z, eps
            return self.draw_minibatch(batch, seed, base + k, counter=counter)

        nxt, tok = draw(0), 0
        for k in range(cfg.n_critic):
            x, z, eps = nxt
            hold = {}

            def ahead(k=k, hold=hold):
                # while the gradients of update k are being all-reduced: inputs + G(z) + interpolates of k+1,
                # or, after the last critic update, the noise + G(z) of the generator update
                if k + 1 < cfg.n_critic:
                    hold["n"] = draw(k + 1)
                    hold["tok"] = self.critic_prepare(*hold["n"])
                elif generator_update:
                    hold["zg"] = self.noise_prior(batch, seed=seed, call_id=base + cfg.n_critic, counter=counter,
                                                  data_max_value=self._data_max_value)
                    hold["tok"] = self.generator_prepare(hold["zg"])
                else:
                    hold["tok"] = 0

            r = self.critic_step(x, z, eps, read=read, overlap=ahead if self.world > 1 else None, token=tok)
            if self.world == 1:
                ahead()
            nxt = hold.get("n")
            zg = hold.get("zg")
            tok = hold["tok"]
            if read:
                out.update(r)
        if generator_update:
            r = self.generator_step(zg, read=read, token=tok)
            if read:
                out.update(r)
        return out if read else None

    def capture_iteration(self, batch: int, seed: int = 0, start_it: int = 0, generator_update: bool = True):
        """Capture one resident-data iteration (every launch of n_critic critic updates + 1 generator
        update, the device RNG, the gather and, in data-parallel runs, the NCCL all-reduces) into a CUDA
        graph and return `replay()`.  The reference pays 4 session calls per iteration [W:184-199]; here a
        replay is one graph launch, which is what makes the reference's own batch 32 fast.  The RNG call-id
        base lives in a device counter advanced inside the graph, so every replay is the next iteration
        (bit-identical to calling train_iteration_resident(it, ...) with it = start_it, start_it + 1, ...)."""
        cfg = self.cfg
        per_it = cfg.n_critic + 1 if cfg.step_order != "gen_first" else 1
        counter = torch.full((1,), start_it * (cfg.n_critic + 1), dtype=torch.int32, device=f"cuda:{self.device}")
        # warm-up outside capture (lazy attribute / occupancy queries, NCCL communicators), then rewind
        side = torch.cuda.Stream(device=f"cuda:{self.device}")
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.train_iteration_resident(0, batch, seed=seed, counter=counter, generator_update=generator_update)
        torch.cuda.current_stream(self.device).wait_stream(side)
        counter += cfg.n_critic + 1            # the eager warm-up WAS iteration start_it; replays continue from there
        torch.cuda.synchronize()
        l0 = self.launch_count
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.train_iteration_resident(0, batch, seed=seed, counter=counter, generator_update=generator_update)
            rc = self._pl.wgan_counter_add(C.c_void_p(counter.data_ptr()), cfg.n_critic + 1, self._st())
            if rc != 0:
                raise _lib.WganError("counter_add: " + self.lib.wgan_last_error(None).decode())
        self._graph_keepalive = (graph, counter)
        self.graph_launches_per_replay = self.launch_count - l0
        _ = per_it
        return graph.replay

    # ------------------------------------------------------------------ training driver [W:166-208]
    def train(self, n_train_steps: Optional[int] = None, batch_size: Optional[int] = None, seed: int = 0,
              logger=None, log_every: int = 100, checkpoint_path: Optional[str] = None, save_every: int = 1000,
              use_graph: bool = True, start_step: int = 0):
        """The reference's training loop [W:182-203] on the resident matrix (set_data): n_train_steps
        iterations (default cfg.n_train_steps = 10000 [W:9]) at cfg.batch_size [W:11]; a state dump every
        `save_every` = 1000 iterations and once at the end [W:201-205]; scalar summaries loss/gen, loss/disc
        (the reference writes them every iteration [W:196-197]; here every `log_every` to avoid a host sync per
        iteration).  Iterations are CUDA-graph replays.  Returns the last logged losses."""
        from .checkpoint import save_checkpoint
        cfg = self.cfg
        n = cfg.n_train_steps if n_train_steps is None else int(n_train_steps)
        B = cfg.batch_size if batch_size is None else int(batch_size)
        assert self._data is not None, "call set_data(cells_by_genes) first"
        last = {}
        replay = None
        for i in range(start_step, start_step + n):
            want_log = (logger is not None or i == start_step + n - 1) and ((i + 1) % log_every == 0 or i == start_step + n - 1)
            if want_log or not use_graph:
                last = self.train_iteration_resident(i, B, seed=seed, read=want_log) or last
                replay = None                       # the eager iteration advanced past the graph's counter
                if want_log and logger is not None:
                    logger.log(i, last)
            else:
                if replay is None:
                    replay = self.capture_iteration(B, seed=seed, start_it=i)   # runs iteration i eagerly
                else:
                    replay()
            if checkpoint_path is not None and (i % save_every == 0 or i == start_step + n - 1):
                # intermediate dumps carry the step in their name like the reference's Saver [W:203] (none of them
                # overwrites another); the dump after the last iteration goes to checkpoint_path itself [W:205]
                save_checkpoint(self, checkpoint_path, global_step=i + 1, step_suffix=i != start_step + n - 1)
        return last

    def generate(self, num_cells: Optional[int] = None, seed: int = 0) -> torch.Tensor:
        """[W:217-221]: num_cells_generate (+1) cells from fresh noise -- one batched pass instead of 501
        single-cell session runs."""
        n = (self.cfg.num_cells_generate + 1) if num_cells is None else int(num_cells)
        return self.sample_cells(n, seed=seed)

    # ------------------------------------------------------------------ bulk sampling (BASELINE config 5)
    def sample_cells(self, n_cells: int, chunk: Optional[int] = None, seed: int = 0, sink=None,
                     first_row: Optional[int] = None) -> Optional[torch.Tensor]:
        """Generate n_cells cells in chunks of <= max_batch rows: z ~ noise_prior on the device, G(z).
        The latent of global cell r is a pure function of (seed, r), so the result is independent of the
        chunking and ranks shard the cell range without communication (rank k takes rows
        [k * n_cells, (k + 1) * n_cells) unless first_row is given).  `sink(row0, cells)` receives every chunk
        (e.g. to copy to host / disk); without a sink the cells are returned as one [n_cells, gex] tensor."""
        chunk = min(chunk or self.max_batch, self.max_batch)
        base = self.rank * n_cells if first_row is None else first_row
        out = None
        if sink is None:
            out = torch.empty((n_cells, self.Gl), device=f"cuda:{self.device}", dtype=torch.float32)
        buf = None if sink is None else torch.empty((chunk, self.Gl), device=f"cuda:{self.device}", dtype=torch.float32)
        for r0 in range(0, n_cells, chunk):
            n = min(chunk, n_cells - r0)
            z = self.noise_prior(n, seed=seed, call_id=0x5A3D0000, row0=base + r0,
                                 data_max_value=self._data_max_value)
            dst = (out[r0:r0 + n] if sink is None else buf[:n])[:, :self.cfg.gex_size]
            self.generator(z, out=dst)
            if sink is not None:
                sink(base + r0, dst)
        return None if sink is not None else out[:, :self.cfg.gex_size]

    # ------------------------------------------------------------------ host-fed iteration (e2e)
    def train_iteration_host(self, xs, zs, epss, z_gen, read: bool = True):
        """One WGAN-GP iteration fed from HOST float32 buffers, the way the reference feeds
        ``sess.run`` [W:196,199]: per critic update x [B, gex], z [B, noise], eps [B]; then z_gen for
        the generator update.  Pinned host tensors are copied on a second stream into two
        alternating device staging sets so the H2D copy of update k+1 overlaps the kernels of update k.
        Returns the losses as host floats (device->host read inside the call)."""
        cfg = self.cfg
        B = xs[0].shape[0]
        dev = f"cuda:{self.device}"
        if getattr(self, "_stage", None) is None or self._stage[0][0].shape[0] != B:
            self._stage = [(torch.empty((B, cfg.gex_size), device=dev), torch.empty((B, cfg.noise_input_size), device=dev),
                            torch.empty((B,), device=dev)) for _ in range(2)]
            self._zgen = torch.empty((B, cfg.noise_input_size), device=dev)
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._free = [torch.cuda.Event() for _ in range(2)]
            self._zgen_free = None         # recorded after the generator step that last read _zgen
            self._nfeed = 0
        main = torch.cuda.current_stream(self.device)
        out = {}
        n = len(xs)
        for k in range(n):
            slot = self._nfeed % 2
            xd, zd, ed = self._stage[slot]
            with torch.cuda.stream(self._copy_stream):
                if self._nfeed >= 2:
                    self._copy_stream.wait_event(self._free[slot])
                xd.copy_(xs[k], non_blocking=True)
                zd.copy_(zs[k], non_blocking=True)
                if cfg.lambda_gp != 0:
                    ed.copy_(epss[k], non_blocking=True)
                if k == n - 1:
                    # the previous iteration's generator step may still be reading the latent (the slot events above
                    # only cover critic updates, which with n_critic <= 2 were recorded BEFORE that generator step)
                    if self._zgen_free is not None:
                        self._copy_stream.wait_event(self._zgen_free)
                    self._zgen.copy_(z_gen, non_blocking=True)
                self._ready[slot].record(self._copy_stream)
            main.wait_event(self._ready[slot])
            r = self.critic_step(xd, zd, ed if cfg.lambda_gp != 0 else None, read=read and k == n - 1)
            self._free[slot].record(main)
            self._nfeed += 1
            if r:
                out.update(r)
        r = self.generator_step(self._zgen, read=read)
        self._zgen_free = torch.cuda.Event()
        self._zgen_free.record(main)
        if r:
            out.update(r)
        return out if read else None

    def train_iteration_host_indices_graph(self, idxs, zs, epss, z_gen, read: bool = True,
                                           generator_update: bool = True):
        """Same contract as train_iteration_host_indices, replayed as ONE CUDA graph: the host->device
        copies out of the caller's pinned buffers, the gathers, all updates and the device->host copy of the
        loss scalars are graph nodes.  The graph is keyed by the addresses of the pinned host buffers (refill
        the same buffers for every iteration, as a feed ring would); new buffers trigger a re-capture."""
        cfg = self.cfg
        key = tuple(t.data_ptr() for t in list(idxs) + list(zs) + list(epss) + [z_gen]) + (bool(generator_update),)
        cache = getattr(self, "_hgraph", None)
        if cache is None or cache[0] != key:
            for t in list(idxs) + list(zs) + list(epss) + [z_gen]:
                assert t.is_pinned(), "graph replay needs pinned host buffers"
            dev = f"cuda:{self.device}"
            B = idxs[0].shape[0]
            n = len(idxs)
            d_idx = [torch.empty((B,), device=dev, dtype=torch.int64) for _ in range(n)]
            d_z = [torch.empty((B, cfg.noise_input_size), device=dev) for _ in range(n + 1)]
            d_eps = [torch.empty((B,), device=dev) for _ in range(n)]
            h_scal = torch.zeros((8,), dtype=torch.float32).pin_memory()

            copy_stream = torch.cuda.Stream(device=dev)
            ready = [torch.cuda.Event() for _ in range(n + 1)]

            def body():
                # all host->device copies of the iteration go to a second stream (a fork inside the graph), so
                # the feed of update k+1 overlaps the kernels of update k
                main = torch.cuda.current_stream(self.device)
                copy_stream.wait_stream(main)
                with torch.cuda.stream(copy_stream):
                    for k in range(n):
                        d_idx[k].copy_(idxs[k], non_blocking=True)
                        d_z[k].copy_(zs[k], non_blocking=True)
                        if cfg.lambda_gp != 0:
                            d_eps[k].copy_(epss[k], non_blocking=True)
                        ready[k].record(copy_stream)
                    d_z[n].copy_(z_gen, non_blocking=True)
                    ready[n].record(copy_stream)

                def inputs(k):             # gather of update k [W:189-191] + the critic-independent half of the step
                    main.wait_event(ready[k])
                    x = self.gather(d_idx[k])
                    e = d_eps[k] if cfg.lambda_gp != 0 else None
                    return x, e, self.critic_prepare(x, d_z[k], e)

                cur, gtok = inputs(0), 0
                for k in range(n):
                    x, e, tok = cur
                    hold = {}

                    def ahead(k=k, hold=hold):     # runs next to the gradient exchange of update k (data-parallel runs)
                        if k + 1 < n:
                            hold["n"] = inputs(k + 1)
                        elif generator_update:
                            main.wait_event(ready[n])
                            hold["g"] = self.generator_prepare(d_z[n])

                    self.critic_step(x, d_z[k], e, read=False, overlap=ahead if self.world > 1 else None, token=tok)
                    if self.world == 1:
                        ahead()
                    cur, gtok = hold.get("n"), hold.get("g", 0)
                if generator_update:
                    self.generator_step(d_z[n], read=False, token=gtok)
                main.wait_event(ready[n])
                main.wait_stream(copy_stream)
                h_scal[0:4].copy_(self._grad_arena[NET_CRITIC][-4:], non_blocking=True)
                h_scal[4:8].copy_(self._grad_arena[NET_GENERATOR][-4:], non_blocking=True)

            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                body()                                   # eager warm-up (this IS a training iteration)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            self._hgraph = (key, g, h_scal, (d_idx, d_z, d_eps))
            first = True
        else:
            first = False
        _, g, h_scal, _ = self._hgraph
        if not first:
            g.replay()
        if not read:
            return None
        torch.cuda.current_stream(self.device).synchronize()
        s = h_scal.numpy()
        wass = float(s[0] - s[1])
        return {"wasserstein": wass, "gp": float(s[2]), "critic_loss": wass + float(s[2]), "gen_loss": -float(s[4])}

    def train_iteration_host_indices(self, idxs, zs, epss, z_gen, read: bool = True):
        """One WGAN-GP iteration on the RESIDENT training matrix (set_data) with the per-update host
        inputs of the reference loop [W:188-193] fed from pinned host memory: the minibatch cell
        indices (np.random.randint), the latent noise and the interpolation weights.  The column
        gather of [W:189-191] happens on the device.  Copies of update k+1 overlap update k."""
        cfg = self.cfg
        B = idxs[0].shape[0]
        dev = f"cuda:{self.device}"
        if getattr(self, "_istage", None) is None or self._istage[0][0].shape[0] != B:
            self._istage = [(torch.empty((B,), device=dev, dtype=torch.int64),
                             torch.empty((B, cfg.noise_input_size), device=dev), torch.empty((B,), device=dev))
                            for _ in range(2)]
            self._izgen = torch.empty((B, cfg.noise_input_size), device=dev)
            self._icopy = torch.cuda.Stream(device=dev)
            self._iready = [torch.cuda.Event() for _ in range(2)]
            self._ifree = [torch.cuda.Event() for _ in range(2)]
            self._izgen_free = None
            self._infeed = 0
        main = torch.cuda.current_stream(self.device)
        out = {}
        n = len(idxs)
        for k in range(n):
            slot = self._infeed % 2
            id_, zd, ed = self._istage[slot]
            with torch.cuda.stream(self._icopy):
                if self._infeed >= 2:
                    self._icopy.wait_event(self._ifree[slot])
                id_.copy_(idxs[k], non_blocking=True)
                zd.copy_(zs[k], non_blocking=True)
                if cfg.lambda_gp != 0:
                    ed.copy_(epss[k], non_blocking=True)
                if k == n - 1:
                    if self._izgen_free is not None:       # see train_iteration_host
                        self._icopy.wait_event(self._izgen_free)
                    self._izgen.copy_(z_gen, non_blocking=True)
                self._iready[slot].record(self._icopy)
            main.wait_event(self._iready[slot])
            x = self.gather(id_)
            r = self.critic_step(x, zd, ed if cfg.lambda_gp != 0 else None, read=read and k == n - 1)
            self._ifree[slot].record(main)
            self._infeed += 1
            if r:
                out.update(r)
        r = self.generator_step(self._izgen, read=read)
        self._izgen_free = torch.cuda.Event()
        self._izgen_free.record(main)
        if r:
            out.update(r)
        return out if read else None

    # ------------------------------------------------------------------ latent fit
    def latent_fit(self, x: torch.Tensor, steps: int, learn_rate: float = 1e-2, seed: int = 0,
                   z0: Optional[torch.Tensor] = None, return_state: bool = False):
        """min_z sum_g (G(z) - x)^2 per cell with G frozen, TF-Adam on z.  Returns (z, loss), or with
        `return_state` (z, loss, m, v): the Adam slots of z (after one step from zero state m = (1 - beta1) dL/dz)."""
        xp, ldx, rows = self._mat(x, self.cfg.gex_size)
        Z = self.cfg.noise_input_size
        z = (z0.clone() if z0 is not None else self.noise_prior(rows, seed=seed, call_id=0x7FFF0000)).contiguous()
        m = torch.zeros_like(z)
        v = torch.zeros_like(z)
        pw = torch.tensor([self.cfg.beta1, self.cfg.beta2, 0, 0], device=z.device, dtype=torch.float32)
        loss = torch.empty((rows,), device=z.device, dtype=torch.float32)
        for _ in range(steps):
            self._ck(self.lib.wgan_latent_fit_step(self._h, C.c_void_p(z.data_ptr()), C.c_void_p(m.data_ptr()),
                                                   C.c_void_p(v.data_ptr()), C.c_void_p(pw.data_ptr()), xp, ldx,
                                                   C.c_void_p(loss.data_ptr()), rows, learn_rate, self._st()),
                     "latent_fit_step")
        return (z, loss, m, v) if return_state else (z, loss)


def generator(trainer: WGANTrainer, z, reuse: bool = False):
    """Functional alias with the reference signature generator(z, reuse) [W:61]."""
    return trainer.generator(z)


def discriminator(trainer: WGANTrainer, x, reuse: bool = False):
    """Functional alias with the reference signature discriminator(x, reuse) [W:103]."""
    return trainer.discriminator(x)

"""State dump / restore and scalar logging around the hot path.

The reference saves every global variable with tf.train.Saver [W:171,201-205]: the 12 layer
variables, their Adam slots, the two optimizers' beta powers and global_step.  TensorFlow's
checkpoint format is not reproducible without TensorFlow, so the same state is written as an
.npz keyed by the TF variable names ("<var>/Adam" = m, "<var>/Adam_1" = v, as TF names the
slots), which is also the mapping needed to import a real TF checkpoint later.
Logging mirrors the reference's scalar summaries loss/gen and loss/disc [W:145-146].
"""
from __future__ import annotations

import os

import numpy as np

from .trainer import PARAM_NAMES, NET_GENERATOR, NET_CRITIC


def _npz_path(path: str) -> str:
    return path if path.endswith(".npz") else path + ".npz"


def save_checkpoint(trainer, path: str, global_step: int = 0, step_suffix: bool = False) -> str:
    """Write the state dump.  Data-parallel replicas are bit-identical, so only rank 0 writes (followed by a barrier
    when torch.distributed is up).  The file is written next to its destination and renamed into place, so a crash
    during a save never destroys the previous checkpoint; `step_suffix` appends -<global_step> like the
    reference's Saver [W:203].  Returns the path written (".npz" is appended when missing, for save AND load)."""
    dest = _npz_path(path)
    if step_suffix:
        dest = dest[:-4] + f"-{int(global_step)}.npz"
    if getattr(trainer, "rank", 0) == 0:
        out = {"global_step": np.int64(global_step)}
        for k in PARAM_NAMES:
            out[k] = trainer.get_tensor(k, "param")
            out[k + "/Adam"] = trainer.get_tensor(k, "m")
            out[k + "/Adam_1"] = trainer.get_tensor(k, "v")
        for net, scope in ((NET_GENERATOR, "generator_opt"), (NET_CRITIC, "critic_opt")):
            b1, b2 = trainer.adam_powers(net)
            out[scope + "/beta1_power"] = np.float32(b1)
            out[scope + "/beta2_power"] = np.float32(b2)
        tmp = dest + f".tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            np.savez(f, **out)
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, dest)
    dist = getattr(trainer, "dist", None)
    if getattr(trainer, "world", 1) > 1 and dist is not None and dist.is_available() and dist.is_initialized():
        dist.barrier(group=getattr(trainer, "group", None))
    return dest


def load_checkpoint(trainer, path: str) -> int:
    ck = np.load(path if os.path.exists(path) else _npz_path(path))
    for k in PARAM_NAMES:
        trainer.set_tensor(k, ck[k], "param")
        trainer.set_tensor(k, ck[k + "/Adam"], "m")
        trainer.set_tensor(k, ck[k + "/Adam_1"], "v")
    for net, scope in ((NET_GENERATOR, "generator_opt"), (NET_CRITIC, "critic_opt")):
        trainer.set_adam_powers(net, float(ck[scope + "/beta1_power"]), float(ck[scope + "/beta2_power"]))
    return int(ck["global_step"])


class ScalarLogger:
    """TensorBoard scalars under the reference's tags [W:145-146] (+ loss/gp); histograms are out of scope."""

    def __init__(self, logdir: str):
        from torch.utils.tensorboard import SummaryWriter
        self.w = SummaryWriter(logdir)

    def log(self, step: int, scalars: dict) -> None:
        if "gen_loss" in scalars:
            self.w.add_scalar("loss/gen", scalars["gen_loss"], step)
        if "critic_loss" in scalars:
            self.w.add_scalar("loss/disc", scalars["critic_loss"], step)
        if "gp" in scalars:
            self.w.add_scalar("loss/gp", scalars["gp"], step)

    def close(self) -> None:
        self.w.close()

"""State dump / restore and scalar logging around the hot path.

The reference saves every global variable with tf.train.Saver [W:171,201-205]: the 12 layer
variables, their Adam slots, the two optimizers' beta powers and global_step.  TensorFlow's
checkpoint format is not reproducible without TensorFlow, so the same state is written as an
.npz keyed by the TF variable names ("<var>/Adam" = m, "<var>/Adam_1" = v, as TF names the
slots), which is also the mapping needed to import a real TF checkpoint later.
Logging mirrors the reference's scalar summaries loss/gen and loss/disc [W:145-146].
"""
from __future__ import annotations

import numpy as np

from .trainer import PARAM_NAMES, NET_GENERATOR, NET_CRITIC


def save_checkpoint(trainer, path: str, global_step: int = 0) -> None:
    out = {"global_step": np.int64(global_step)}
    for k in PARAM_NAMES:
        out[k] = trainer.get_tensor(k, "param")
        out[k + "/Adam"] = trainer.get_tensor(k, "m")
        out[k + "/Adam_1"] = trainer.get_tensor(k, "v")
    for net, scope in ((NET_GENERATOR, "generator_opt"), (NET_CRITIC, "critic_opt")):
        b1, b2 = trainer.adam_powers(net)
        out[scope + "/beta1_power"] = np.float32(b1)
        out[scope + "/beta2_power"] = np.float32(b2)
    np.savez(path, **out)


def load_checkpoint(trainer, path: str) -> int:
    ck = np.load(path)
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

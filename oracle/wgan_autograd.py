"""Independent fp64 torch.autograd implementation of the same graph (TEST INFRASTRUCTURE ONLY).

Used only to cross-check the hand-derived numpy oracle (oracle/wgan_oracle.py): forward through
torch ops, every gradient through autograd, the gradient penalty through
autograd-of-autograd (create_graph=True).  Follows [W:61-83], [W:103-124], [W:137-138].
"""
import numpy as np
import torch
import torch.nn.functional as F

from .wgan_oracle import GEN_NAMES, CRITIC_NAMES


def _t(P, names, requires_grad):
    return [torch.tensor(np.asarray(P[k], dtype=np.float64), requires_grad=requires_grad)
            for k in names]


def _gen(gp, z, a):
    h = F.leaky_relu(z @ gp[0] + gp[1], a)
    h = F.leaky_relu(h @ gp[2] + gp[3], a)
    return F.leaky_relu(h @ gp[4] + gp[5], a)


def _critic(cp, x, a):
    h = F.leaky_relu(x @ cp[0] + cp[1], a)
    h = F.leaky_relu(h @ cp[2] + cp[3], a)
    return h @ cp[4] + cp[5]


def critic_step_grads(P, x, z, eps, lam, alpha=0.2):
    gp_ = _t(P, GEN_NAMES, False)
    cp = _t(P, CRITIC_NAMES, True)
    x = torch.tensor(np.asarray(x, np.float64))
    z = torch.tensor(np.asarray(z, np.float64))
    xt = _gen(gp_, z, alpha).detach()
    wass = _critic(cp, xt, alpha).mean() - _critic(cp, x, alpha).mean()
    pen = torch.zeros((), dtype=torch.float64)
    if lam != 0.0:
        e = torch.tensor(np.asarray(eps, np.float64)).reshape(-1, 1)
        xh = (e * x + (1 - e) * xt).requires_grad_(True)
        (gx,) = torch.autograd.grad(_critic(cp, xh, alpha).sum(), xh, create_graph=True)
        pen = lam * ((gx.norm(dim=1) - 1) ** 2).mean()
    loss = wass + pen
    grads = torch.autograd.grad(loss, cp, allow_unused=True)
    out = {k: (g.numpy() if g is not None else np.zeros_like(P[k]))
           for k, g in zip(CRITIC_NAMES, grads)}
    return {"wasserstein": wass.item(), "gp": pen.item(), "critic_loss": loss.item()}, out


def generator_step_grads(P, z, alpha=0.2):
    gp_ = _t(P, GEN_NAMES, True)
    cp = _t(P, CRITIC_NAMES, False)
    z = torch.tensor(np.asarray(z, np.float64))
    loss = -_critic(cp, _gen(gp_, z, alpha), alpha).mean()
    grads = torch.autograd.grad(loss, gp_)
    return {"gen_loss": loss.item()}, {k: g.numpy() for k, g in zip(GEN_NAMES, grads)}


def latent_fit_grads(P, z, x, alpha=0.2):
    gp_ = _t(P, GEN_NAMES, False)
    z = torch.tensor(np.asarray(z, np.float64), requires_grad=True)
    x = torch.tensor(np.asarray(x, np.float64))
    loss = ((_gen(gp_, z, alpha) - x) ** 2).sum(dim=1)
    (dz,) = torch.autograd.grad(loss.sum(), z)
    return loss.detach().numpy(), dz.numpy()

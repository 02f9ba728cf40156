"""Gaussian parameter-exploring policy gradients (PGPE) run next to the backprop planner
(planner.py:1677-2143 of the reference; loop integration planner.py:3270-3296).

PGPE is a pure consumer of the exact forward rollout kernel: each step samples plan parameters
theta = mu +- eps (and, with super-symmetric sampling, mu +- eps*), evaluates their exact
test returns (4 forward launches of ``rk_rollout_forward`` through the planner's test path), and
moves (mu, sigma) along the baseline-free gradient estimates with two Adam optimisers.  The
parameter vectors are the planner's flat plan vectors (T*A words for a straight-line plan, the
MLP weights for a DRP), so everything around the rollouts is a handful of elementwise device ops
on a few thousand words; no host synchronisation happens inside ``update``.
"""
from __future__ import annotations

import math
from typing import Any, Dict, Optional, Tuple

import torch


class PGPE:
    """Base class for all PGPE strategies (planner.py:1684)."""


class GaussianPGPE(PGPE):
    """PGPE with a Gaussian parameter distribution (planner.py:1714)."""

    def __init__(self, batch_size: int = 1, steps_per_update: int = 1, init_sigma: float = 1.0,
                 sigma_range: Tuple[float, float] = (1e-5, 1e5), scale_reward: bool = True,
                 min_reward_scale: float = 1e-5, super_symmetric: bool = True,
                 super_symmetric_accurate: bool = True, project_samples: bool = False,
                 optimizer: Any = "adam", optimizer_kwargs_mu: Optional[Dict] = None,
                 optimizer_kwargs_sigma: Optional[Dict] = None,
                 clip_grad_mu: Optional[float] = None, clip_grad_sigma: Optional[float] = None,
                 ema_decay: Optional[float] = None, start_entropy_coeff: float = 1e-2,
                 end_entropy_coeff: float = 1e-5, max_kl_update: Optional[float] = None,
                 eps: float = 1e-10) -> None:
        name = optimizer if isinstance(optimizer, str) else getattr(optimizer, "__name__", "")
        if name not in ("adam", "rmsprop", "sgd"):
            raise NotImplementedError(f"GaussianPGPE: optimizer <{name}> (adam, rmsprop, sgd)")
        self.optimizer_name = name
        self.project_samples = bool(project_samples)       # planner.py:1927-1940
        self.ema_decay = None if ema_decay is None else float(ema_decay)   # optax.ema link, P:1790-1792
        self.batch_size = batch_size
        self.steps_per_update = steps_per_update
        self.init_sigma = init_sigma
        self.sigma_range = sigma_range
        self.scale_reward = scale_reward
        self.min_reward_scale = min_reward_scale
        self.super_symmetric = super_symmetric
        self.super_symmetric_accurate = super_symmetric_accurate
        self.clip_grad_mu = clip_grad_mu
        self.clip_grad_sigma = clip_grad_sigma
        self.start_entropy_coeff = start_entropy_coeff
        self.end_entropy_coeff = end_entropy_coeff
        self.optimizer_kwargs_mu = optimizer_kwargs_mu or {"learning_rate": 0.1}
        self.optimizer_kwargs_sigma = optimizer_kwargs_sigma or {"learning_rate": 0.1}
        self.max_kl = max_kl_update
        self.eps = eps

    def __str__(self) -> str:
        return (f"[INFO] PGPE hyper-parameters:\n"
                f"    method        ={type(self).__name__}\n"
                f"    batch_size    ={self.batch_size}\n"
                f"    init_sigma    ={self.init_sigma}\n"
                f"    sigma_range   ={self.sigma_range}\n"
                f"    scale_reward  ={self.scale_reward}\n"
                f"    super_symmetric={self.super_symmetric} (accurate={self.super_symmetric_accurate})\n"
                f"    max_kl_update ={self.max_kl}\n")

    # ---------------------------------------------------------------------------------
    def compile(self, planner) -> None:
        self.pl = planner
        dev, P, n = planner.device, planner.parallel_updates, planner.n_params
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)   # noqa: E731
        self.mu, self.sigma = z(P, n), z(P, n)
        self.adam = {k: (z(P, n), z(P, n)) for k in ("mu", "sigma")}
        self.ema = {k: z(P, n) for k in ("mu", "sigma")}
        self.r_max = torch.full((P,), -math.inf, device=dev)
        self.loss4 = z(4)
        self.step = 0
        self.cand = z(n)
        if self.start_entropy_coeff == 0:
            self.decay = 0.0
        else:
            self.decay = (self.end_entropy_coeff / self.start_entropy_coeff) ** 0.01

    def initialize(self, params: torch.Tensor, generator: torch.Generator) -> None:
        """planner.py:1856-1871: mu = plan parameters, sigma = init_sigma."""
        self.mu.copy_(params)
        self.sigma.fill_(self.init_sigma)
        for m, v in self.adam.values():
            m.zero_()
            v.zero_()
        for e in self.ema.values():
            e.zero_()
        self.r_max.fill_(-math.inf)
        self.step = 0
        self.gen = generator

    def policy_params(self) -> torch.Tensor:
        return self.mu

    # ---- planner.py:1888-1916 -----------------------------------------------------------
    def eps_star(self, sigma: torch.Tensor, epsilon: torch.Tensor) -> torch.Tensor:
        phi = 0.67449 * sigma
        a = (sigma - epsilon.abs()) / sigma
        if self.super_symmetric_accurate:
            aa = a.abs()
            c1, c2, c3 = -0.06655, -0.9706, 0.124
            term_neg = torch.exp(torch.where(
                aa == 1., torch.full_like(aa, 2. * c1 + c2),
                c1 * aa * (aa * aa - 1.) / torch.log(aa) + c2 * aa))
            safe = torch.clamp(1. - aa ** 3, min=torch.finfo(aa.dtype).tiny)
            term_pos = torch.exp(aa) / torch.pow(safe, c3 * aa)
            return torch.where(epsilon == 0., torch.zeros_like(epsilon),
                               torch.sign(epsilon) * phi * torch.where(a < 0, term_neg, term_pos))
        return torch.sign(epsilon) * phi * torch.exp(a)

    # ---- planner.py:1954-2003 -----------------------------------------------------------
    def grads(self, epsilon, epsilon_star, sigma, r, m, ent):
        r1, r2, r3, r4 = r[0], r[1], r[2], r[3]
        mrs = self.min_reward_scale
        if self.super_symmetric:
            if self.scale_reward:
                s1 = torch.clamp(m - (r1 + r2) / 2, min=mrs)
                s2 = torch.clamp(m - (r3 + r4) / 2, min=mrs)
            else:
                s1 = s2 = torch.ones_like(r1)
            g_mu = -(((r1 - r2) / (2 * s1)) * epsilon + ((r3 - r4) / (2 * s2)) * epsilon_star)
            eps_tau = torch.where(r1 + r2 >= r3 + r4, epsilon, epsilon_star)
            s = eps_tau.square() / sigma.pow(3) - 1. / sigma
            sc = torch.clamp(m - (r1 + r2 + r3 + r4) / 4, min=mrs) if self.scale_reward \
                else torch.ones_like(r1)
            r_sigma = ((r1 + r2) - (r3 + r4)) / (4 * sc)
        else:
            sc = torch.clamp(m - (r1 + r2) / 2, min=mrs) if self.scale_reward else torch.ones_like(r1)
            g_mu = -((r1 - r2) / (2 * sc)) * epsilon
            s = epsilon.square() / sigma.pow(3) - 1. / sigma
            sc2 = torch.clamp(m.abs(), min=mrs) if self.scale_reward else torch.ones_like(r1)
            r_sigma = (r1 + r2) / (2 * sc2)
        g_sigma = -(r_sigma * s + ent / sigma)
        return g_mu, g_sigma

    def _adam(self, which: str, p: int, x: torch.Tensor, g: torch.Tensor, kw: Dict, clip) -> torch.Tensor:
        """One step of the meta-optimiser chain of planner.py:1786-1792 on the mean or the deviations:
        [clip_by_global_norm] -> adam (bias-corrected, step counted from 1) / rmsprop / sgd -> [ema]."""
        if clip is not None:
            g = g * torch.clamp(clip / torch.linalg.vector_norm(g).clamp_min(1e-30), max=1.0)
        m, v = self.adam[which][0][p], self.adam[which][1][p]
        t = self.step
        if self.optimizer_name == "adam":
            b1, b2, eps = kw.get("b1", 0.9), kw.get("b2", 0.999), kw.get("eps", 1e-8)
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            mh, vh = m / (1 - b1 ** t), v / (1 - b2 ** t)
            upd = mh / (vh.sqrt() + eps)
        elif self.optimizer_name == "rmsprop":
            d, eps = kw.get("decay", 0.9), kw.get("eps", 1e-8)
            v.mul_(d).addcmul_(g, g, value=1 - d)
            upd = g / torch.sqrt(v + eps)
        else:
            mom = kw.get("momentum", None)
            if mom:
                m.mul_(mom).add_(g)
                upd = m.clone()
            else:
                upd = g
        delta = -kw.get("learning_rate", 0.1) * upd
        if self.ema_decay is not None:             # optax.ema with debias
            e = self.ema[which][p]
            e.mul_(self.ema_decay).add_(delta, alpha=1 - self.ema_decay)
            delta = e / (1 - self.ema_decay ** t)
        return x + delta

    # ---- planner.py:2059-2128 -----------------------------------------------------------
    def update(self, progress: float) -> None:
        pl = self.pl
        ent = self.start_entropy_coeff * (self.decay ** progress)
        lo, hi = self.sigma_range
        for _ in range(self.steps_per_update):
            self.step += 1
            for p in range(pl.parallel_updates):
                mu, sigma = self.mu[p], self.sigma[p]
                g_mu, g_sigma = torch.zeros_like(mu), torch.zeros_like(sigma)
                # planner.py:2045-2058: every batch member scales its gradient by the running maximum
                # as it stood BEFORE the batch joined with its own four rewards; the maxima are merged
                # after the batch (vmap over the members, then jnp.max)
                r_before, r_after = self.r_max[p].clone(), self.r_max[p].clone()
                for _b in range(self.batch_size):
                    epsilon = sigma * torch.randn(mu.shape, device=mu.device, generator=self.gen)
                    es = self.eps_star(sigma, epsilon) if self.super_symmetric else epsilon
                    cands = (mu + epsilon, mu - epsilon) + \
                        ((mu + es, mu - es) if self.super_symmetric else ())
                    for i, c in enumerate(cands):
                        self.cand.copy_(c)
                        if self.project_samples:      # planner.py:1927-1940
                            pl.project_params(self.cand, p)
                        pl._test_kernels(p, params=self.cand, scratch=True,
                                         loss_out=self.loss4[i:i + 1])
                    if not self.super_symmetric:
                        self.loss4[2:4].copy_(self.loss4[0:2])
                    r = -self.loss4
                    m_b = torch.maximum(r_before, r.max())
                    r_after = torch.maximum(r_after, m_b)
                    gm, gs = self.grads(epsilon, es, sigma, r, m_b, ent)
                    g_mu += gm / self.batch_size
                    g_sigma += gs / self.batch_size
                self.r_max[p:p + 1].copy_(r_after.reshape(1))
                old_mu, old_sigma = mu.clone(), sigma.clone()
                state = {k: (m[p].clone(), v[p].clone()) for k, (m, v) in self.adam.items()}
                ema_state = {k: e[p].clone() for k, e in self.ema.items()}
                new_mu = self._adam("mu", p, mu, g_mu, self.optimizer_kwargs_mu, self.clip_grad_mu)
                new_sigma = self._adam("sigma", p, sigma, g_sigma, self.optimizer_kwargs_sigma,
                                       self.clip_grad_sigma).clamp(lo, hi)
                if self.max_kl is not None:       # planner.py:2085-2096
                    kl = 0.5 * torch.sum(2 * torch.log(new_sigma / old_sigma)
                                         + (old_sigma / new_sigma).square()
                                         + ((new_mu - old_mu) / new_sigma).square() - 1)
                    scale = torch.clamp(torch.sqrt(self.max_kl / (kl + self.eps)), max=1.0)
                    for k, (m, v) in self.adam.items():
                        m[p].copy_(state[k][0])
                        v[p].copy_(state[k][1])
                        self.ema[k][p].copy_(ema_state[k])
                    new_mu = self._adam("mu", p, old_mu, g_mu * scale, self.optimizer_kwargs_mu,
                                        self.clip_grad_mu)
                    new_sigma = self._adam("sigma", p, old_sigma, g_sigma * scale,
                                           self.optimizer_kwargs_sigma,
                                           self.clip_grad_sigma).clamp(lo, hi)
                mu.copy_(new_mu)
                sigma.copy_(new_sigma)
                pl.project_params(mu, p)          # plan.projection (planner.py:2098-2100)

"""CPU ORACLE (test infrastructure) -- PARITY UNPINNED -- planner side.

Restates, in torch-CPU fp32 with autograd:
  * straight-line plan policies      pyRDDLGym_jax/core/planner.py:741-868
  * box / concurrency projections    planner.py:489-617, 878-959
  * batched scan over the horizon    compiler.py:551-619
  * return, loss                     planner.py:2747-2822
  * optimiser updates                planner.py:2349-2387, 2885-2899 (optax semantics assumed:
    rmsprop  nu <- d nu + (1-d) g^2, update = g * rsqrt(nu + eps);  adam with bias
    correction, step counted from 1;  clip_by_global_norm;  apply_updates p + u;
    optax.contrib.reduce_on_plateau as a plain state machine)
  * deep reactive policy             planner.py:1027-1119 (GumbelSoftmaxTopK, FiLM, time
    embeddings), 1223-1530 (input layer with StaticNormalizer / LayerNorm, hidden layers and
    activations, logit / mu / log_sigma / dense_topk heads, train and test policies)
"""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from .model_eval import OracleModel, REAL, INT, safe_log


# ---- straight-line plan ----------------------------------------------------------------
def bound_action(lo, hi, param: torch.Tensor) -> torch.Tensor:
    """_jax_bound_action, planner.py:620-628."""
    lo = torch.as_tensor(np.asarray(lo, np.float32))
    hi = torch.as_tensor(np.asarray(hi, np.float32))
    fl, fh = torch.isfinite(lo), torch.isfinite(hi)
    lo0, hi0 = torch.where(fl, lo, torch.zeros(())), torch.where(fh, hi, torch.zeros(()))
    sp = torch.nn.functional.softplus
    both = lo0 + (hi0 - lo0) * torch.sigmoid(param)
    out = torch.where(fl & fh, both, torch.where(fl, lo0 + sp(param),
                                                 torch.where(fh, hi0 - sp(-param), param)))
    return out


def slp_pooled_softmax(model, params: Dict[str, torch.Tensor], t: int, weight: float,
                       gumbel=None):
    """planner.py:764-808: softmax over the pooled boolean logits of step t, unstacked per fluent
    (1 - p for fluents whose no-op value is true).  gumbel: per fluent [B, n] draws added to the
    logits of a stochastic plan (planner.py:801-805); the result then has a leading batch axis."""
    names = [v for v in model.action_fluents if model.variable_ranges[v] == "bool"]
    z = torch.cat([params[v][t].reshape(-1) for v in names])
    if gumbel is not None:
        z = z.unsqueeze(0) + torch.cat([x.reshape(x.shape[0], -1) for x in gumbel], dim=1)
    out = torch.softmax(weight * z, dim=-1)
    acts, pos = {}, 0
    for v in names:
        n = params[v][t].numel()
        a = out[..., pos:pos + n]
        pos += n
        noop = bool(np.ravel(np.asarray(model.action_fluents[v]))[0])
        acts[v] = (1.0 - a) if noop else a
    return acts


def slp_policy_noise_spec(model, wrap_sigmoid=True, action_noise="normal", pooled_softmax=False):
    """Draw sites of a stochastic straight-line plan, in the order planner.py:811-833 splits
    its key: (kind, shape) per action fluent that receives noise.  The pooled '__bool__' entry
    (one Gumbel draw over all boolean parameters, planner.py:801-803) comes first and is listed
    fluent by fluent."""
    spec = []
    if pooled_softmax:
        spec += [("gumbel", tuple(model.variable_shape(v))) for v in model.action_fluents
                 if model.variable_ranges[v] == "bool"]
    for name in model.action_fluents:
        shape = tuple(model.variable_shape(name))
        if model.variable_ranges[name] == "bool":
            if pooled_softmax:
                continue
            if wrap_sigmoid:
                spec.append(("logistic", shape))
        else:
            spec.append((action_noise, shape))
    return spec


def slp_entropy(model, params: Dict[str, torch.Tensor], t: int, wrap_sigmoid=True,
                sigmoid_weight=1.0, sigma_range=(1e-5, 1e3), sigma_entropy_grad=False,
                pooled_softmax=False, softmax_weight=1.0):
    """Entropy term of one step of a stochastic plan (planner.py:797-800, 812-830)."""
    from oracle.model_eval import safe_log
    ent = torch.zeros(())
    if pooled_softmax:
        z = torch.cat([params[v][t].reshape(-1) for v in model.action_fluents
                       if model.variable_ranges[v] == "bool"])
        log_prob = torch.log_softmax(softmax_weight * z, dim=0)
        ent = ent - torch.sum(torch.exp(log_prob) * log_prob)
    for name in model.action_fluents:
        p = params[name][t]
        if model.variable_ranges[name] == "bool":
            if wrap_sigmoid and not pooled_softmax:
                w = sigmoid_weight[name] if isinstance(sigmoid_weight, dict) else sigmoid_weight
                prob = torch.sigmoid(w * p)
                ent = ent - torch.sum(prob * safe_log(prob) + (1. - prob) * safe_log(1. - prob))
        else:
            ls = torch.clamp(params[name + "/log_sigma"][t], float(np.log(sigma_range[0])),
                             float(np.log(sigma_range[1])))
            ent = ent + torch.sum(ls if sigma_entropy_grad else ls.detach())
    return ent


def slp_train_actions(model, params: Dict[str, torch.Tensor], t: int, wrap_sigmoid=True,
                      sigmoid_weight: Dict[str, float] | float = 1.0, wrap_non_bool=False,
                      bounds=None, stochastic=False, sigma_range=(1e-5, 1e3), pnoise=None,
                      pooled_softmax=False, softmax_weight=1.0):
    """planner.py:786-838.  ``pnoise``: this step's policy draws [B, *shape] in the order of
    slp_policy_noise_spec (stochastic plans only)."""
    actions = {}
    k = 0
    pooled = {}
    if pooled_softmax:
        gumbel = None
        if stochastic:
            k = sum(1 for v in model.action_fluents if model.variable_ranges[v] == "bool")
            gumbel = pnoise[:k]
        pooled = slp_pooled_softmax(model, params, t, softmax_weight, gumbel)
    for name in model.action_fluents:
        p = params[name][t]
        if name in pooled:
            actions[name] = pooled[name] if pooled[name].dim() > 1 else pooled[name].unsqueeze(0)
            continue
        if model.variable_ranges[name] == "bool":
            w = sigmoid_weight[name] if isinstance(sigmoid_weight, dict) else sigmoid_weight
            if stochastic and wrap_sigmoid:                            # planner.py:811-817
                p = p.reshape(-1).unsqueeze(0) + pnoise[k].reshape(pnoise[k].shape[0], -1)
                k += 1
            a = torch.sigmoid(w * p) if wrap_sigmoid else p            # planner.py:741-745
        else:
            a = p
            if stochastic:                                             # planner.py:825-833
                ls = torch.clamp(params[name + "/log_sigma"][t], float(np.log(sigma_range[0])),
                                 float(np.log(sigma_range[1])))
                z = pnoise[k].reshape(pnoise[k].shape[0], -1)
                k += 1
                a = a.reshape(-1).unsqueeze(0) + torch.exp(ls).reshape(-1).unsqueeze(0) * z
            if wrap_non_bool:
                a = bound_action(np.asarray(bounds[name][0]).reshape(-1),
                                 np.asarray(bounds[name][1]).reshape(-1), a)   # planner.py:755-762
        actions[name] = a if a.dim() > 1 and stochastic else a.unsqueeze(0)
    return actions


def slp_test_actions(model, params: Dict[str, torch.Tensor], t: int, bounds, wrap_sigmoid=True,
                     wrap_non_bool=False, pooled_softmax=False, softmax_weight=1.0):
    """planner.py:842-868."""
    actions = {}
    thr = 0.0 if wrap_sigmoid else 0.5
    pooled = slp_pooled_softmax(model, params, t, softmax_weight) if pooled_softmax else {}
    for name in model.action_fluents:
        p = params[name][t]
        prange = model.variable_ranges[name]
        if name in pooled:
            a = pooled[name] > 0.5
        elif prange == "bool":
            a = p > thr
        else:
            lo, hi = bounds[name]
            lo = torch.as_tensor(np.asarray(lo, np.float32))
            hi = torch.as_tensor(np.asarray(hi, np.float32))
            if wrap_non_bool:
                p = bound_action(lo, hi, p)
            a = torch.minimum(torch.maximum(p, lo), hi)
            if prange != "real":
                a = torch.round(a).to(INT)
        actions[name] = a.unsqueeze(0)
    return actions


def expand_actions(actions, B):
    return {k: v.expand((B,) + tuple(v.shape[1:])) for k, v in actions.items()}


# ---- rollouts ------------------------------------------------------------------------------
def rollout(om: OracleModel, policy: Callable[[int, Dict[str, torch.Tensor]], Dict[str, torch.Tensor]],
            B: int, T: int, noise_steps: List[List[torch.Tensor]], subs=None, log_fluents=()):
    """compiler.py:596-619: scan of the batched policy step.  Returns dict with reward [B,T],
    error [B,T], final fluents and optional per-step fluent logs [B,T,...]."""
    fls, nfls = om.init_fls_nfls(B, subs)
    rewards, errs = [], []
    logs = {f: [] for f in log_fluents}
    states = []
    spec = None
    for t in range(T):
        states.append({s: fls[s] for s in om.model.state_fluents})
        n_pol = getattr(policy, "n_noise", 0)      # stochastic plans draw before the model does
        step_noise = noise_steps[t] if noise_steps else None
        if n_pol:
            actions = expand_actions(policy(t, fls, step_noise[:n_pol]), B)
            step_noise = step_noise[n_pol:]
        else:
            actions = expand_actions(policy(t, fls), B)
        if om.relaxed:
            actions = {k: v.to(REAL) for k, v in actions.items()}
        fls, r, e, spec = om.step(fls, nfls, actions, step_noise, B)
        rewards.append(r)
        errs.append(e)
        for f in log_fluents:
            logs[f].append(fls[f])
    out = {"reward": torch.stack(rewards, 1) if T else torch.zeros(B, 0),
           "error": torch.stack(errs, 1) if T else torch.zeros(B, 0, dtype=torch.int64),
           "final": {s: fls[s] for s in om.model.state_fluents},
           "states": states, "noise_spec": spec,
           "fluents": {f: torch.stack(v, 1) for f, v in logs.items()}}
    return out


def returns_from_rewards(rewards: torch.Tensor, gamma: float = 1.0):
    """planner.py:2751-2764."""
    if gamma != 1:
        T = rewards.shape[-1]
        disc = torch.pow(torch.tensor(gamma, dtype=REAL), torch.arange(T, dtype=REAL))
        rewards = rewards * disc.unsqueeze(0)
    return rewards.sum(1)


def mean_loss(rewards: torch.Tensor, gamma: float = 1.0):
    """planner.py:2817-2818 with utility='mean'."""
    return -returns_from_rewards(rewards, gamma).mean()


# ---- optimisers (optax semantics) ---------------------------------------------------------------
def clip_by_global_norm(grads: List[torch.Tensor], max_norm: float):
    gn = torch.sqrt(sum((g * g).sum() for g in grads))
    if gn > max_norm:
        return [g * (max_norm / gn) for g in grads]
    return grads


def rmsprop_step(p, g, nu, lr, decay=0.9, eps=1e-8):
    nu = decay * nu + (1. - decay) * g * g
    return p - lr * g / torch.sqrt(nu + eps), nu


def adam_step(p, g, mu, nu, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam.  The moment updates take (1 - decay) from the Python floats (double precision,
    rounded once); optax.bias_correction forms ``1 - decay ** count`` with ``count`` an int32 ARRAY,
    i.e. a float32 power of the float32-rounded decay."""
    mu = b1 * mu + (1. - b1) * g
    nu = b2 * nu + (1. - b2) * g * g
    t = torch.tensor(float(step), dtype=REAL)
    mh = mu / (1. - torch.pow(torch.tensor(b1, dtype=REAL), t))
    nh = nu / (1. - torch.pow(torch.tensor(b2, dtype=REAL), t))
    return p - lr * mh / (torch.sqrt(nh) + eps), mu, nu


def sgd_step(p, g, lr):
    return p - lr * g


def sgd_momentum_step(p, g, trace, lr, momentum):
    """optax.sgd(momentum=m): trace <- g + m * trace ; update = -lr * trace."""
    trace = g + momentum * trace
    return p - lr * trace, trace


# ---- projections ------------------------------------------------------------------------------
def bool_box(wrap_sigmoid: bool, min_action_prob: float, weight: float):
    """planner.py:878-882: clip bounds of a boolean action parameter."""
    if wrap_sigmoid:
        logit = lambda a: math.log(a / (1. - a))      # noqa: E731
        return logit(min_action_prob) / weight, logit(1. - min_action_prob) / weight
    return min_action_prob, 1. - min_action_prob


def project_box(model, params, bounds, wrap_sigmoid=True, min_action_prob=1e-6, weight=1.0):
    """planner.py:884-904."""
    out = {}
    for name, p in params.items():
        if model.variable_ranges[name] == "bool":
            lo, hi = bool_box(wrap_sigmoid, min_action_prob, weight)
            out[name] = torch.clamp(p, lo, hi)
        else:
            lo, hi = bounds[name]
            lo = torch.as_tensor(np.asarray(lo, np.float32))
            hi = torch.as_tensor(np.asarray(hi, np.float32))
            out[name] = torch.minimum(torch.maximum(p, lo), hi)
    return out


def sogbofa_project_row(actions: Dict[str, torch.Tensor], noop: Dict[str, bool], allowed: int,
                        max_iter: int, min_action: float, max_action: float):
    """planner.py:540-616 on one time step: actions -> adjusted actions (before the
    action->parameter inverse map).  Returns (new_actions, converged)."""
    def surplus_k(acts):
        total, k = torch.tensor(0., dtype=REAL), 0
        for var, a in acts.items():
            if noop[var]:
                a = 1 - a
            total = total + a.sum()
            k = k + int(torch.count_nonzero(a))
        return torch.clamp(total - allowed, min=0.), k

    acts = dict(actions)
    surplus, k = surplus_k(acts)
    it = 0
    while it < max_iter and surplus > 0 and k > 0:
        amount = surplus / k
        acts = {var: (torch.clamp(a + amount, max=1.) if noop[var] else torch.clamp(a - amount, min=0.))
                for var, a in acts.items()}
        surplus, k = surplus_k(acts)
        it += 1
    converged = not bool(surplus > 0)
    total_bool = 0
    for var, a in acts.items():
        total_bool += int(torch.count_nonzero(a < 0.5 if noop[var] else a > 0.5))
    excess = max(total_bool - allowed, 0)
    new = {}
    for var, a in acts.items():
        a = torch.clamp(a, min_action, max_action)
        flat = a.reshape(-1).clone()
        active = (flat < 0.5) if noop[var] else (flat > 0.5)
        ranks = torch.cumsum(active.to(torch.int64), 0)
        replace = active & (ranks <= excess)
        flat[replace] = 0.5
        new[var] = flat.reshape(a.shape)
        excess = max(excess - int(torch.count_nonzero(replace)), 0)
    return new, converged


def sorting_project_row(logits: Dict[str, torch.Tensor], noop: Dict[str, bool], allowed: int,
                        wrap_sigmoid: bool, bool_threshold: float, box: Dict[str, Tuple[float, float]]):
    """planner.py:499-527 on one time step: shift every boolean parameter by the surplus of
    the (allowed+1)-th largest score over the threshold, then clip to the boolean box."""
    scores = []
    for var, p in logits.items():
        flat = p.reshape(-1)
        if noop[var]:
            flat = -flat if wrap_sigmoid else 1. - flat
        scores.append(flat)
    scores = torch.cat(scores)
    descending = torch.sort(scores, descending=True).values
    surplus = torch.clamp(descending[allowed] - bool_threshold, min=0.)
    out = {}
    for var, p in logits.items():
        shifted = p + surplus if noop[var] else p - surplus
        out[var] = torch.clamp(shifted, box[var][0], box[var][1])
    return out, True


def project_max_constraint(model, params: Dict[str, torch.Tensor], method: str, allowed: int,
                           wrap_sigmoid=True, min_action_prob=1e-6, weight=1.0, max_iter=100):
    """planner.py:947-954: the per-step projection vmapped over the horizon axis; ``params``
    are the box-projected plan parameters {var: [T, n]} (non-boolean entries pass through)."""
    bools = [v for v in params if model.variable_ranges[v] == "bool"]
    noop = {v: bool(np.ravel(model.action_fluents[v])[0]) for v in bools}
    T = next(iter(params.values())).shape[0]
    out = {v: p.clone() for v, p in params.items()}
    converged = True
    thr = 0.0 if wrap_sigmoid else 0.5
    box = {v: bool_box(wrap_sigmoid, min_action_prob, weight) for v in bools}
    for t in range(T):
        row = {v: params[v][t] for v in bools}
        if method == "sorting":
            new, ok = sorting_project_row(row, noop, allowed, wrap_sigmoid, thr, box)
        else:
            acts = {v: (torch.sigmoid(weight * p) if wrap_sigmoid else p) for v, p in row.items()}
            acts, ok = sogbofa_project_row(acts, noop, allowed, max_iter, min_action_prob,
                                           1. - min_action_prob)
            new = {v: (torch.log(a / (1. - a)) / weight if wrap_sigmoid else a)
                   for v, a in acts.items()}
        converged = converged and ok
        for v in bools:
            out[v][t] = new[v]
    return out, converged


# ---- deep reactive policy (planner.py:1122-1534) ---------------------------------------------
def drp_input_order(model) -> List[str]:
    """Order in which observed fluents enter the state vector (planner.py:1302-1313): the
    fluent dict crosses a jit boundary, so it iterates in sorted-key order; non-bool fluents
    first, then bool ones."""
    # partially observed models: the observ-fluents only (their carried copies in the observation view)
    names = sorted(getattr(model, "observed_inputs", None) or model.state_fluents)
    return [v for v in names if model.variable_ranges[v] != "bool"] + \
           [v for v in names if model.variable_ranges[v] == "bool"]


def drp_norm_layers(model, normalize: bool, normalize_per_layer: bool):
    """The LayerNorm modules of the input layer (planner.py:1259-1282, 1300-1319) as a list of
    (flax name, [fluents]); falls back like the reference when a normalised tensor has size 1."""
    if not normalize:
        return []
    non_bool = [v for v in drp_input_order(model) if model.variable_ranges[v] != "bool"]
    sizes = [model.variable_size(v) for v in non_bool]
    if normalize_per_layer and any(n == 1 for n in sizes):
        normalize_per_layer = False
    if normalize_per_layer:
        return [("norm_" + v.replace("-", "_"), [v]) for v in non_bool]
    if sum(sizes) <= 1:
        return []
    return [("norm", non_bool)]


def static_normalizer_transform(fls, stats):
    """StaticNormalizer.transform (planner.py:326-338): (values - lower) / (upper - lower) on the
    fluents that have stats, the others unchanged."""
    out = {}
    for var, values in fls.items():
        if stats is not None and var in stats:
            lower = torch.as_tensor(np.asarray(stats[var][0], np.float32))
            upper = torch.as_tensor(np.asarray(stats[var][1], np.float32))
            out[var] = (values.to(REAL) - lower) / (upper - lower)
        else:
            out[var] = values
    return out


def drp_state_vector(model, params, fls, normalize: bool = False,
                     normalize_per_layer: bool = False, eps: float = 1e-6, preprocess=None):
    """DRPStateLayer (planner.py:1287-1320): ravel + concatenate the observed fluents (non-bool
    first), LayerNorm on the non-bool part (flax nn.LayerNorm: biased variance E[x^2] - E[x]^2,
    epsilon 1e-6, learned scale and bias under params['inputs'][name])."""
    B = next(iter(fls.values())).shape[0]
    if preprocess is not None:          # planner.py:1293-1296: before everything else
        fls = static_normalizer_transform(fls, preprocess)
    cols = {v: fls[v].reshape(B, -1).to(REAL) for v in drp_input_order(model)}
    for name, members in drp_norm_layers(model, normalize, normalize_per_layer):
        x = torch.cat([cols[v] for v in members], dim=1)
        mean = x.mean(1, keepdim=True)
        var = torch.clamp((x * x).mean(1, keepdim=True) - mean * mean, min=0.0)
        layer = params["inputs"][name]
        y = (x - mean) * torch.rsqrt(var + eps) * layer["scale"] + layer["bias"]
        o = 0
        for v in members:
            n = cols[v].shape[1]
            cols[v] = y[:, o:o + n]
            o += n
    return torch.cat([cols[v] for v in drp_input_order(model)], dim=1)


def drp_init(model, topology, generator: torch.Generator, stochastic: bool = True,
             normalize: bool = False, normalize_per_layer: bool = False):
    """variance_scaling(2.0, 'fan_in', 'truncated_normal') kernels, zero biases
    (planner.py:1129-1130; flax nn.Dense), as a dict layer -> {'kernel' [in, out], 'bias'};
    LayerNorm scales start at one and biases at zero (flax defaults)."""
    def kernel(fan_in, fan_out):
        std = math.sqrt(2.0 / fan_in) / 0.87962566103423978
        w = torch.empty(fan_in, fan_out)
        torch.nn.init.trunc_normal_(w, mean=0., std=1., a=-2., b=2., generator=generator)
        return w * std
    D = sum(model.variable_size(v) for v in drp_input_order(model))
    params, fan = {}, D
    for name, members in drp_norm_layers(model, normalize, normalize_per_layer):
        n = sum(model.variable_size(v) for v in members)
        params.setdefault("inputs", {})[name] = {"scale": torch.ones(n), "bias": torch.zeros(n)}
    for i, h in enumerate(topology):
        params[f"dense_{i}"] = {"kernel": kernel(fan, h), "bias": torch.zeros(h)}
        fan = h
    n_bool = sum(model.variable_size(v) for v in model.action_fluents
                 if model.variable_ranges[v] == "bool")
    pooled = model.max_allowed_actions < n_bool
    for var in model.action_fluents:
        n = model.variable_size(var)
        name = var.replace("-", "_")
        if model.variable_ranges[var] == "bool" and pooled:
            continue
        if model.variable_ranges[var] == "bool":
            params[f"dense_{name}_logit"] = {"kernel": kernel(fan, n), "bias": torch.zeros(n)}
        else:
            params[f"dense_{name}_mu"] = {"kernel": kernel(fan, n), "bias": torch.zeros(n)}
            if stochastic:
                params[f"dense_{name}_log_sigma"] = {"kernel": kernel(fan, n),
                                                     "bias": torch.zeros(n)}
    if pooled:
        params["dense_topk"] = {"kernel": kernel(fan, n_bool), "bias": torch.zeros(n_bool)}
    return params


def drp_actions(model, params, fls, noise: Optional[Dict[str, torch.Tensor]], train: float,
                bounds, topology_len: int, stochastic: bool = True,
                sigma_range: Tuple[float, float] = (1e-5, 1e3), sigmoid_weight: float = 1.0,
                wrap_non_bool: bool = False, normalize: bool = False,
                normalize_per_layer: bool = False, norm_eps: float = 1e-6,
                softmax_weight: float = 1.0, sigma_entropy_grad: bool = False,
                step: int = 0, time_embedding: Optional[str] = None, time_dim: int = 8,
                activation: str = "tanh", preprocess=None):
    """planner.py:1342-1449 + 1468-1476 for one decision epoch.  fls: name -> [B, *objs];
    noise: action name -> [B, n] standard draws (Gumbel for boolean fluents under a binding
    max-nondef-actions, Logistic for other boolean fluents).  Returns (actions name ->
    [B, *objs], entropy [B])."""
    B = next(iter(fls.values())).shape[0]
    x = drp_state_vector(model, params, fls, normalize, normalize_per_layer, norm_eps, preprocess)
    h = x
    t_emb = None if time_embedding is None else drp_time_embedding(params, step, time_embedding,
                                                                   time_dim)
    for i in range(topology_len):
        layer = params[f"dense_{i}"]
        h = DRP_ACTIVATIONS[activation](h @ layer["kernel"] + layer["bias"])
        if t_emb is not None:           # FiLM (planner.py:1083-1092, 1362-1363)
            film = params[f"film_{i}"]
            gamma = t_emb @ film["gamma"]["kernel"] + film["gamma"]["bias"]
            beta = t_emb @ film["beta"]["kernel"] + film["beta"]["bias"]
            h = gamma * h + beta
    ls_lo, ls_hi = math.log(sigma_range[0]), math.log(sigma_range[1])
    actions, entropy = {}, torch.zeros(B, dtype=REAL)
    bools = [v for v in model.action_fluents if model.variable_ranges[v] == "bool"]
    n_bool = sum(model.variable_size(v) for v in bools)
    pooled = model.max_allowed_actions < n_bool             # planner.py:1322-1326
    for var in model.action_fluents:
        name = var.replace("-", "_")
        shape = (B,) + tuple(model.variable_shape(var))
        if model.variable_ranges[var] == "bool":
            if pooled:
                continue
            layer = params[f"dense_{name}_logit"]
            logits = h @ layer["kernel"] + layer["bias"]
            if stochastic:
                prob = torch.sigmoid(sigmoid_weight * logits)
                entropy = entropy - (prob * safe_log(prob) + (1. - prob) * safe_log(1. - prob)).sum(1)
                logits = logits + train * noise[var]
            actions[var] = torch.sigmoid(sigmoid_weight * logits).reshape(shape)
            continue
        layer = params[f"dense_{name}_mu"]
        a = h @ layer["kernel"] + layer["bias"]
        if stochastic:
            layer = params[f"dense_{name}_log_sigma"]
            ls = torch.clamp(h @ layer["kernel"] + layer["bias"], ls_lo, ls_hi)
            entropy = entropy + (ls if sigma_entropy_grad else ls.detach()).sum(1)
            a = a + train * torch.exp(ls) * noise[var]
        lo, hi = bounds[var]
        lo = torch.as_tensor(np.asarray(lo, np.float32)).reshape(1, -1)
        hi = torch.as_tensor(np.asarray(hi, np.float32)).reshape(1, -1)
        if wrap_non_bool:                                     # planner.py:1411-1415
            a = bound_action(lo, hi, a)
        else:
            a = torch.minimum(torch.maximum(a, lo), hi)       # planner.py:1472-1475
        actions[var] = a.reshape(shape)
    if pooled:      # one 'dense_topk' head + GumbelSoftmaxTopK (planner.py:1435-1447, 1027-1080)
        layer = params["dense_topk"]
        logits = h @ layer["kernel"] + layer["bias"]
        if stochastic:
            log_prob = torch.log_softmax(softmax_weight * logits, dim=1)
            entropy = entropy - (torch.exp(log_prob) * log_prob).sum(1)
        g = torch.cat([noise[v].reshape(B, -1) for v in bools], dim=1) if stochastic else 0.0
        out = gumbel_softmax_topk(logits, g, int(model.max_allowed_actions), softmax_weight, train)
        o = 0
        for v in bools:                                        # planner.py:1453-1465
            n = model.variable_size(v)
            a = out[:, o:o + n]
            if bool(np.ravel(np.asarray(model.action_fluents[v]))[0]):
                a = 1.0 - a
            actions[v] = a.reshape((B,) + tuple(model.variable_shape(v)))
            o += n
    return actions, entropy


# jax.nn functions a .cfg can name for the hidden layers (planner.py:144-150, 1358-1361)
DRP_ACTIVATIONS = {
    "tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid,
    "softplus": torch.nn.functional.softplus, "elu": torch.nn.functional.elu,
    "leaky_relu": lambda x: torch.nn.functional.leaky_relu(x, 0.01),
    # jax.nn.gelu defaults to the tanh approximation; jax.nn.silu = jax.nn.swish
    "gelu": lambda x: torch.nn.functional.gelu(x, approximate="tanh"),
    "silu": torch.nn.functional.silu, "swish": torch.nn.functional.silu}


def drp_time_embedding(params, step: int, kind: str, dim: int = 8):
    """SinusoidalTimeEmbedding (planner.py:1108-1119): angles t * 10000^(-2 i / dim), (sin, cos)
    pairs interleaved and cut to the first `dim` entries; FixedTimeEmbedding (planner.py:1095-1105):
    row `step` of a learned [max_t + 1, dim] table."""
    if kind == "fixed":
        return params["fixed_time_embed"]["kernel"][step]
    freqs = 10000.0 ** (-2.0 * torch.arange(dim, dtype=REAL) / dim)
    angles = float(step) * freqs
    return torch.stack([torch.sin(angles), torch.cos(angles)], dim=-1).reshape(-1)[:dim]


def drp_film_init(params, topology, generator: torch.Generator, kind: str, dim: int = 8,
                  max_t: int = 500):
    """flax defaults: Dense kernels lecun_normal (variance_scaling(1, fan_in, truncated normal)),
    zero biases; the fixed table normal(stddev 0.01)."""
    def kernel(fan_in, fan_out):
        std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
        w = torch.empty(fan_in, fan_out)
        torch.nn.init.trunc_normal_(w, mean=0., std=1., a=-2., b=2., generator=generator)
        return w * std
    for i, h in enumerate(topology):
        params[f"film_{i}"] = {"gamma": {"kernel": kernel(dim, h), "bias": torch.zeros(h)},
                               "beta": {"kernel": kernel(dim, h), "bias": torch.zeros(h)}}
    if kind == "fixed":
        params["fixed_time_embed"] = {"kernel": torch.randn(max_t + 1, dim, generator=generator) * 0.01}
    return params


def gumbel_softmax_topk(logits, gumbel, allowed: int, weight: float, train: float):
    """GumbelSoftmaxTopK.__call__ (planner.py:1034-1080).  logits [B, n]; gumbel: Gumbel(0, 1)
    draws of the same shape (0 = not stochastic)."""
    if allowed == 1:                                           # planner.py:1039-1045
        return torch.softmax(logits + train * gumbel, dim=1)
    if train:                                                  # soft_top_k, planner.py:1048-1063
        l = logits + gumbel
        khot = torch.zeros_like(logits)
        for _ in range(allowed):
            l = l + safe_log(1.0 - khot)
            khot = khot + torch.softmax(weight * l, dim=1)
        return khot
    l = logits.clone()                                         # hard_top_k, planner.py:1066-1078
    khot = torch.zeros_like(logits)
    rows = torch.arange(logits.shape[0])
    for _ in range(allowed):
        idx = torch.argmax(l, dim=1)
        l[rows, idx] = -float("inf")
        khot[rows, idx] = 1.0
    return khot


def drp_test_actions(model, params, fls, bounds, topology_len: int, stochastic: bool = True,
                     sigmoid_weight: float = 1.0, wrap_non_bool: bool = False,
                     normalize: bool = False, normalize_per_layer: bool = False,
                     norm_eps: float = 1e-6, softmax_weight: float = 1.0, step: int = 0,
                     time_embedding: Optional[str] = None, time_dim: int = 8,
                     activation: str = "tanh", preprocess=None):
    """planner.py:1482-1499: train = 0, then threshold / round / clip to the fluent's range."""
    zeros = {v: torch.zeros(next(iter(fls.values())).shape[0], model.variable_size(v))
             for v in model.action_fluents}
    acts, _ = drp_actions(model, params, fls, zeros, 0.0, bounds, topology_len, stochastic,
                          sigmoid_weight=sigmoid_weight, wrap_non_bool=wrap_non_bool,
                          normalize=normalize, normalize_per_layer=normalize_per_layer,
                          norm_eps=norm_eps, softmax_weight=softmax_weight, step=step,
                          time_embedding=time_embedding, time_dim=time_dim, activation=activation,
                          preprocess=preprocess)
    out = {}
    for var, a in acts.items():
        prange = model.variable_ranges[var]
        if prange == "bool":
            out[var] = a > 0.5
        elif prange == "int" or prange in model.enum_types:
            lo, hi = bounds[var]
            a = torch.minimum(torch.maximum(a, torch.as_tensor(np.asarray(lo, np.float32))),
                              torch.as_tensor(np.asarray(hi, np.float32)))
            out[var] = torch.round(a).to(INT)
        else:
            lo, hi = bounds[var]
            out[var] = torch.minimum(torch.maximum(a, torch.as_tensor(np.asarray(lo, np.float32))),
                                     torch.as_tensor(np.asarray(hi, np.float32)))
    return out


# ---- optax.contrib.reduce_on_plateau (planner.py:2375-2384, 2889-2891) --------------------------
def reduce_on_plateau_init():
    """optax (>= 0.2.7, un-vendored third-party code: parity unpinned) ReduceLROnPlateauState."""
    return {"scale": 1.0, "best_value": math.inf, "plateau_count": 0, "cooldown_count": 0,
            "count": 0, "avg_value": 0.0}


def reduce_on_plateau_step(state, value: float, factor=0.1, patience=10, rtol=1e-4, atol=0.0,
                           cooldown=0, accumulation_size=1, min_scale=0.0):
    """One `update` of optax.contrib.reduce_on_plateau as a plain Python state machine; returns the
    new state, whose 'scale' multiplies the optimiser's updates of this step."""
    count = state["count"] + 1
    avg = (state["count"] * state["avg_value"] + value) / count
    if count != accumulation_size:
        return dict(state, count=count, avg_value=avg)
    improved = avg < (1.0 - rtol) * state["best_value"] - atol
    best = avg if improved else state["best_value"]
    cur = 0 if improved else state["plateau_count"] + 1
    if state["cooldown_count"] > 0:         # bad steps are ignored while cooling down
        plat, scale, cool = 0, state["scale"], state["cooldown_count"] - 1
    elif cur == patience:
        plat, scale, cool = 0, max(state["scale"] * factor, min_scale), cooldown
    else:
        plat, scale, cool = cur, state["scale"], 0
    return {"scale": scale, "best_value": best, "plateau_count": plat, "cooldown_count": cool,
            "count": 0, "avg_value": 0.0}

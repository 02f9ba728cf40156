"""Packing between the reference's pytrees of [B, *objs] tensors and the kernels' flat,
fluent-major buffers (see include/rk.h), plus generation of the supplied-noise tensor.

The reference keeps fluents in a dict name -> [B, *objs] (JaxRDDLSimState.fls,
compiler.py:54-62) and draws noise with jax.random inside the traced graph.  Here noise is an
explicit input: one slot per draw site, in the order the lowering met
them, which is the order the reference splits its PRNG key (compiler.py:364-366, 1792).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

_GAMMA_CONC: Dict[Any, torch.Tensor] = {}
_TORCH_DT = {"f": torch.float32, "i": torch.int32, "b": torch.int32}


def pack_fluent_major(layout: Dict[str, Tuple[int, int, str]], values: Dict[str, torch.Tensor],
                      B: int, device=None) -> torch.Tensor:
    """dict name -> [B, *objs]  ==>  flat [W*B] words (float32 storage; ints bit-cast)."""
    W = sum(n for _, n, _ in layout.values())
    out = torch.empty(W * B, dtype=torch.float32, device=device)
    for name, (off, n, dt) in layout.items():
        v = values[name].reshape(B, n).to(device)
        if dt == "f":
            v = v.to(torch.float32)
        else:
            v = v.to(torch.int32).view(torch.float32)
        out[off * B:(off + n) * B] = v.reshape(-1)
    return out


def unpack_fluent_major(layout: Dict[str, Tuple[int, int, str]], flat: torch.Tensor, B: int,
                        shapes: Optional[Dict[str, Tuple[int, ...]]] = None,
                        lead: Tuple[int, ...] = ()) -> Dict[str, torch.Tensor]:
    """flat [*lead, W*B] ==> dict name -> [*lead, B, *objs] (typed views, no copy)."""
    W = sum(n for _, n, _ in layout.values())
    flat = flat.reshape(*lead, W * B)
    out = {}
    for name, (off, n, dt) in layout.items():
        v = flat[..., off * B:(off + n) * B]
        if dt != "f":
            v = v.view(torch.int32)
            if dt == "b":
                v = v != 0
        shape = shapes[name] if shapes is not None else (n,)
        out[name] = v.reshape(*lead, B, *shape)
    return out


def draw_noise(noise_layout: Sequence[Tuple[str, Tuple[int, ...], int, int, Any]], T: int, B: int,
               generator: Optional[torch.Generator] = None, device=None) -> torch.Tensor:
    """Supplied-noise tensor [T, N*B] (fluent-major per step) for a program's draw sites."""
    N = sum(n for _, _, n, _, _ in noise_layout)
    out = torch.empty(T, N * B, dtype=torch.float32, device=device)
    tiny = torch.finfo(torch.float32).tiny
    for kind, shape, n, off, extra in noise_layout:
        size = (T, B, n)
        if kind == "normal":
            z = torch.randn(size, generator=generator, device=device)
        else:
            u = torch.rand(size, generator=generator, device=device)
            if kind == "uniform":
                z = u
            elif kind == "exponential":
                z = -torch.log1p(-u)
            elif kind == "gumbel":
                z = -torch.log(-torch.log(u.clamp_min(tiny)))
            elif kind == "logistic":
                u = u.clamp(tiny, 1 - 1e-7)
                z = torch.log(u) - torch.log1p(-u)
            elif kind == "laplace":
                u2 = u - 0.5
                z = -torch.sign(u2) * torch.log1p(-2 * u2.abs())
            elif kind == "cauchy":
                z = torch.tan(np.pi * (u - 0.5))
            elif kind == "gamma":
                conc = torch.as_tensor(np.asarray(extra, dtype=np.float32).reshape(-1),
                                       device="cpu").clamp_min(1e-6)
                if device is not None and torch.device(device).type == "cuda":
                    # standard Gamma(shape) draws on the device from the caller's generator; the
                    # concentrations are uploaded once (no host copy inside a captured graph)
                    key = (conc.numpy().tobytes(), T, B, str(device))
                    a = _GAMMA_CONC.get(key)
                    if a is None:
                        a = conc.to(device).expand(T, B, conc.numel()).contiguous()
                        _GAMMA_CONC[key] = a
                    z = torch._standard_gamma(a, generator=generator)
                else:
                    g = torch.distributions.Gamma(conc, torch.ones_like(conc))
                    # torch's sampler ignores `generator`: seed from it for reproducibility
                    seed = int(torch.randint(0, 2**31 - 1, (1,), generator=generator,
                                             device=device).item()) if generator is not None else None
                    if seed is not None:
                        state = torch.random.get_rng_state()
                        torch.manual_seed(seed)
                    z = g.sample((T, B)).to(device)
                    if seed is not None:
                        torch.random.set_rng_state(state)
            else:
                raise NotImplementedError(f"noise kind {kind}")
        out[:, off * B:(off + n) * B] = z.reshape(T, B * n)
    return out


def noise_as_list(noise_layout, noise: torch.Tensor, t: int, B: int) -> List[torch.Tensor]:
    """The per-draw tensors [B, *shape] of step t (the form the CPU oracle consumes)."""
    out = []
    for kind, shape, n, off, extra in noise_layout:
        out.append(noise[t, off * B:(off + n) * B].reshape((B,) + tuple(shape)))
    return out

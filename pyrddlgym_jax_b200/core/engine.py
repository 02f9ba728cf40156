"""ctypes binding of the C ABI in include/rk.h (csrc/librk.so) + torch buffer plumbing.

PyTorch is used only for device memory, streams and (in parallel.py) torch.distributed; all
compute on the hot path happens in the CUDA kernels behind this binding.  There is no CPU
fallback: using an engine without the native library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np
import torch

from .program import Program

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc", "librk.so")

EXPORTS = [
    "rk_program_create", "rk_program_destroy", "rk_program_query", "rk_program_num_warps",
    "rk_program_attach_cubin", "rk_program_attach_cubin_ex",
    "rk_rollout_forward", "rk_rollout_backward", "rk_reduce_partials", "rk_optimizer_step",
    "rk_mean_loss", "rk_project_slp", "rk_sgemm", "rk_drp_actions_forward",
    "rk_drp_actions_backward", "rk_tf32_split", "rk_tcgemm", "rk_drp_layer0", "rk_colsum",
    "rk_rowdot", "rk_tcgemm_tn", "rk_tile_state", "rk_drp_head_backward",
    "rk_drp_head_backward_ctas", "rk_drp_layer0_backward", "rk_peer_allreduce",
    "rk_state_layernorm_ctas", "rk_state_layernorm_forward", "rk_state_layernorm_backward",
    "rk_drp_topk_limits", "rk_drp_topk_forward", "rk_drp_topk_backward",
    "rk_film_chunks", "rk_film_forward", "rk_film_backward",
    "rk_state_scale_forward", "rk_state_scale_backward", "rk_drp_mlp2_forward",
    "rk_drp_mlp2_backward", "rk_drp_mlp2_image_floats", "rk_drp_mlp2_prepare", "rk_drp_mlp2_set_timing",
    "rk_drp_wgrad_ctas", "rk_drp_wgrad_hidden", "rk_drp_wgrad_heads", "rk_drp_wgrad_input",
    "rk_draw_noise", "rk_noise_advance", "rk_drp_returns_loss", "rk_transpose_split",
    "rk_utility_loss", "rk_optimizer_chain_step", "rk_utility_loss_shard",
    "rk_utility_sorted_work_bytes", "rk_utility_sorted_loss", "rk_peer_allreduce_gather",
    "rk_act_apply", "rk_reduce_optimizer_step",
]


class RkError(RuntimeError):
    pass


class RkSizes(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "kind", "relaxed", "S", "A", "N", "n_mparams", "rpw", "lanes", "warp_words",
        "const_words", "n_prologue", "n_step", "n_epilogue", "warps_per_cta", "smem_bytes",
        "static_err", "tape_words")]


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """Load librk.so and declare every prototype of include/rk.h."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = os.path.abspath(path or LIB_PATH)
    if not os.path.exists(p):
        raise RkError(f"native library {p} is missing: run __graft_entry__.build() "
                      f"(python -m pyrddlgym_jax_b200.build); there is no CPU fallback")
    lib = ctypes.CDLL(p)
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p
    fl = ctypes.c_float
    lib.rk_program_create.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.rk_program_create.restype = ci
    lib.rk_program_destroy.argtypes = [vp]
    lib.rk_program_destroy.restype = None
    lib.rk_program_query.argtypes = [vp, ctypes.POINTER(RkSizes)]
    lib.rk_program_query.restype = ci
    lib.rk_program_attach_cubin.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_char_p]
    lib.rk_program_attach_cubin.restype = ci
    lib.rk_program_attach_cubin_ex.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_char_p, ci, ci, ci]
    lib.rk_program_attach_cubin_ex.restype = ci
    lib.rk_program_num_warps.argtypes = [vp, ci]
    lib.rk_program_num_warps.restype = ci
    lib.rk_rollout_forward.argtypes = [vp, vp, ci, ci] + [cf] * 12
    lib.rk_rollout_forward.restype = ci
    lib.rk_rollout_backward.argtypes = [vp, vp, ci, ci] + [cf] * 11
    lib.rk_rollout_backward.restype = ci
    lib.rk_reduce_partials.argtypes = [vp, cf, ci, ci, cf]
    lib.rk_reduce_partials.restype = ci
    lib.rk_optimizer_step.argtypes = [vp, ci, cf, ci, cf, cf, cf, cf, cf, cf]
    lib.rk_optimizer_step.restype = ci
    lib.rk_optimizer_chain_step.argtypes = [vp, ci, cf, ci, cf, cf, cf, cf, cf, cf, vp, cf, cf, cf, cf,
                                            ctypes.c_uint64, cf]
    lib.rk_optimizer_chain_step.restype = ci
    lib.rk_utility_loss_shard.argtypes = [vp, ci, ci, ci, ci, cf, fl, fl, fl, cf, cf]
    lib.rk_utility_loss_shard.restype = ci
    lib.rk_utility_sorted_work_bytes.argtypes = [ci]
    lib.rk_utility_sorted_work_bytes.restype = ctypes.c_size_t
    lib.rk_utility_sorted_loss.argtypes = [vp, ci, ci, ci, ci, cf, fl, ci, fl, cf, cf, vp, ctypes.c_size_t]
    lib.rk_utility_sorted_loss.restype = ci
    lib.rk_mean_loss.argtypes = [vp, cf, ci, cf, cf]
    lib.rk_mean_loss.restype = ci
    lib.rk_project_slp.argtypes = [vp, ci, ci, ci, cf, ci, vp, vp, vp, vp, ci, ci, ci,
                                   ctypes.c_float, cf]
    lib.rk_project_slp.restype = ci
    fl = ctypes.c_float
    lib.rk_sgemm.argtypes = [vp, ci, ci, ci, ci, cf, ci, cf, ci, cf, ci, cf, cf, ci, cf, ci]
    lib.rk_sgemm.restype = ci
    lib.rk_drp_actions_forward.argtypes = [vp, ci, ci, ci, ci, vp, vp, vp, cf, cf, fl, cf, cf, fl,
                                           fl, fl, ci, cf, cf]
    lib.rk_drp_actions_forward.restype = ci
    lib.rk_drp_actions_backward.argtypes = [vp, ci, ci, ci, ci, vp, vp, vp, cf, cf, fl, cf, cf, fl,
                                            fl, fl, cf, fl, ci, cf, cf]
    lib.rk_drp_actions_backward.restype = ci
    lib.rk_drp_layer0.argtypes = [vp, ci, ci, ci, ci, vp, vp, cf, cf, cf, cf, cf, ci]
    lib.rk_drp_layer0.restype = ci
    lib.rk_drp_head_backward_ctas.argtypes = [ci]
    lib.rk_drp_head_backward_ctas.restype = ci
    lib.rk_drp_head_backward.argtypes = [vp, ci, ci, ci, cf, cf, cf, cf, cf, cf, ci]
    lib.rk_drp_head_backward.restype = ci
    lib.rk_drp_layer0_backward.argtypes = [vp, ci, ci, ci, ci, vp, vp, cf, cf, cf, cf, cf, cf]
    lib.rk_drp_layer0_backward.restype = ci
    lib.rk_peer_allreduce.argtypes = [vp, ci, ci, vp, vp, ci, ci, cf, cf, cf]
    lib.rk_peer_allreduce.restype = ci
    lib.rk_peer_allreduce_gather.argtypes = [vp, ci, ci, vp, vp, ci, ci, cf, cf, cf, ci, vp, vp, cf]
    lib.rk_peer_allreduce_gather.restype = ci
    lib.rk_act_apply.argtypes = [vp, ci, ctypes.c_size_t, cf, cf]
    lib.rk_act_apply.restype = ci
    lib.rk_reduce_optimizer_step.argtypes = [vp, cf, ci, ci, cf, ci, cf, cf, cf, cf, cf, cf, cf]
    lib.rk_reduce_optimizer_step.restype = ci
    lib.rk_state_layernorm_ctas.argtypes = [ci]
    lib.rk_state_layernorm_ctas.restype = ci
    lib.rk_state_layernorm_forward.argtypes = [vp, ci, ci, ci, vp, vp, vp, fl, cf, cf, cf, cf]
    lib.rk_state_layernorm_forward.restype = ci
    lib.rk_state_layernorm_backward.argtypes = [vp, ci, ci, ci, vp, vp, vp, fl, cf, cf, cf, cf, cf, cf]
    lib.rk_state_layernorm_backward.restype = ci
    lib.rk_state_scale_forward.argtypes = [vp, ci, ci, ci, vp, vp, cf, cf, cf, cf]
    lib.rk_state_scale_forward.restype = ci
    lib.rk_state_scale_backward.argtypes = [vp, ci, ci, ci, vp, vp, cf, cf, cf, cf]
    lib.rk_state_scale_backward.restype = ci
    lib.rk_film_chunks.argtypes = [ci]
    lib.rk_film_chunks.restype = ci
    lib.rk_film_forward.argtypes = [vp, ci, ci, cf, cf, cf, cf]
    lib.rk_film_forward.restype = ci
    lib.rk_film_backward.argtypes = [vp, ci, ci, cf, cf, cf, cf, cf, cf, ci]
    lib.rk_film_backward.restype = ci
    lib.rk_drp_topk_limits.argtypes = [vp, vp]
    lib.rk_drp_topk_limits.restype = ci
    lib.rk_drp_topk_forward.argtypes = [vp, ci, ci, ci, ci, vp, vp, vp, cf, cf, ci, ci, fl, ci, cf, cf]
    lib.rk_drp_topk_forward.restype = ci
    lib.rk_drp_topk_backward.argtypes = [vp, ci, ci, ci, ci, vp, vp, vp, cf, cf, ci, fl, cf, fl, cf, cf]
    lib.rk_drp_topk_backward.restype = ci
    lib.rk_colsum.argtypes = [vp, ci, ci, cf, ci, cf, ci, cf, ci]
    lib.rk_colsum.restype = ci
    lib.rk_rowdot.argtypes = [vp, ci, ci, ci, ci, cf, ci, cf, ci, cf, ci, cf]
    lib.rk_rowdot.restype = ci
    lib.rk_tcgemm_tn.argtypes = [vp, ci, ci, ci, ci, cf, ci, cf, ci, cf, cf, ci]
    lib.rk_tcgemm_tn.restype = ci
    lib.rk_tile_state.argtypes = [vp, ci, ci, vp, vp, cf, cf]
    lib.rk_tile_state.restype = ci
    lib.rk_tf32_split.argtypes = [vp, cf, cf, cf, ctypes.c_size_t]
    lib.rk_tf32_split.restype = ci
    lib.rk_tcgemm.argtypes = [vp, ci, ci, ci, ci, cf, ci, cf, cf, ci, cf, ci, cf, cf, ci]
    lib.rk_tcgemm.restype = ci
    lib.rk_drp_mlp2_forward.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, vp, vp] + [cf] * 10
    lib.rk_drp_mlp2_forward.restype = ci
    lib.rk_drp_mlp2_backward.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, vp, vp] + [cf] * 9
    lib.rk_drp_mlp2_backward.restype = ci
    lib.rk_drp_mlp2_image_floats.argtypes = [ci] * 5
    lib.rk_drp_mlp2_image_floats.restype = ci
    lib.rk_drp_mlp2_prepare.argtypes = [vp, ci, ci, ci, ci, cf, cf, cf, cf]
    lib.rk_drp_mlp2_prepare.restype = ci
    lib.rk_draw_noise.argtypes = [vp, ci, ci, ci, vp, vp, ctypes.c_ulonglong, vp, cf, cf]
    lib.rk_draw_noise.restype = ci
    lib.rk_drp_returns_loss.argtypes = [vp, ci, ci, cf, cf, cf, cf, cf, cf, cf]
    lib.rk_drp_returns_loss.restype = ci
    fl = ctypes.c_float
    lib.rk_utility_loss.argtypes = [vp, ci, ci, cf, fl, fl, fl, cf, cf]
    lib.rk_utility_loss.restype = ci
    lib.rk_transpose_split.argtypes = [vp, ci, ci, cf, cf, cf, cf]
    lib.rk_transpose_split.restype = ci
    lib.rk_noise_advance.argtypes = [vp, vp]
    lib.rk_noise_advance.restype = ci
    lib.rk_drp_mlp2_set_timing.argtypes = [vp]
    lib.rk_drp_mlp2_set_timing.restype = ci
    lib.rk_drp_wgrad_ctas.argtypes = []
    lib.rk_drp_wgrad_ctas.restype = ci
    ll = ctypes.c_longlong
    lib.rk_drp_wgrad_hidden.argtypes = [vp, ci, ci, ci, ll, cf, cf, cf, cf, cf]
    lib.rk_drp_wgrad_hidden.restype = ci
    lib.rk_drp_wgrad_heads.argtypes = [vp, ci, ci, ci, ll, cf, cf, cf, cf, cf]
    lib.rk_drp_wgrad_heads.restype = ci
    lib.rk_drp_wgrad_input.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, vp, vp, cf, cf, cf, cf, cf]
    lib.rk_drp_wgrad_input.restype = ci
    if path is None:
        _LIB = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        kind = "invalid argument/program" if rc < 0 else "CUDA error"
        raise RkError(f"{what} failed: {kind} {rc}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RkError("rk kernels take CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise RkError("rk kernels take contiguous tensors")
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class DeviceProgram:
    """A Program uploaded to the current CUDA device."""

    def __init__(self, program: Program, specialize: Optional[bool] = None):
        """``specialize``: attach the per-program generated kernel (default: on, unless the
        environment sets RK_SPECIALIZE=0); otherwise the generic interpreter kernel runs."""
        if specialize is None:
            specialize = os.environ.get("RK_SPECIALIZE", "1") != "0"
        if not torch.cuda.is_available():
            raise RkError("no CUDA device: the rollout kernels have no CPU fallback")
        self.lib = load_library()
        self.program = program
        self.handle = ctypes.c_void_p()
        blob = program.blob
        buf = ctypes.create_string_buffer(blob, len(blob))
        _check(self.lib.rk_program_create(buf, len(blob), ctypes.byref(self.handle)),
               "rk_program_create")
        self.sizes = RkSizes()
        _check(self.lib.rk_program_query(self.handle, ctypes.byref(self.sizes)), "rk_program_query")
        self.specialized = False
        if specialize:
            from .codegen import compile_cubin, KERNEL_NAME
            cubin, _ = compile_cubin(program)
            with open(cubin, "rb") as f:
                image = f.read()
            self._image = ctypes.create_string_buffer(image, len(image))
            if program.lanes == 1:      # thread-per-rollout kernel: registers only (codegen_tpr.py)
                from .codegen_tpr import THREADS
                _check(self.lib.rk_program_attach_cubin_ex(
                    self.handle, self._image, len(image), KERNEL_NAME.encode(), THREADS, 0,
                    int(os.environ.get("RK_CARVEOUT_TPR", "0"))), "rk_program_attach_cubin_ex")
            else:
                _check(self.lib.rk_program_attach_cubin(self.handle, self._image, len(image),
                                                        KERNEL_NAME.encode()),
                       "rk_program_attach_cubin")
            self.specialized = True

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.rk_program_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def num_warps(self, B: int) -> int:
        return self.lib.rk_program_num_warps(self.handle, B)

    def forward(self, B, T, s0, policy=None, actions=None, noise=None, mparams=None, disc=None,
                reward=None, returns=None, err=None, tape=None, final=None, traj=None):
        _check(self.lib.rk_rollout_forward(
            self.handle, _stream(), B, T, _ptr(s0), _ptr(policy), _ptr(actions), _ptr(noise),
            _ptr(mparams), _ptr(disc), _ptr(reward), _ptr(returns), _ptr(err), _ptr(tape),
            _ptr(final), _ptr(traj)), "rk_rollout_forward")

    def backward(self, B, T, tape, gret, policy=None, actions=None, noise=None, mparams=None,
                 disc=None, gradp=None, gact=None, gs0=None, gs_in=None):
        _check(self.lib.rk_rollout_backward(
            self.handle, _stream(), B, T, _ptr(tape), _ptr(policy), _ptr(actions), _ptr(noise),
            _ptr(mparams), _ptr(disc), _ptr(gret), _ptr(gradp), _ptr(gact), _ptr(gs0),
            _ptr(gs_in)), "rk_rollout_backward")


def reduce_partials(gradp: torch.Tensor, rows: int, n: int, grad: torch.Tensor):
    _check(load_library().rk_reduce_partials(_stream(), _ptr(gradp), rows, n, _ptr(grad)),
           "rk_reduce_partials")


OPT_KIND = {"sgd": 0, "rmsprop": 1, "adam": 2}


def optimizer_step(kind: str, hyper: torch.Tensor, n: int, params, state, grad, lo=None, hi=None,
                   zero_flag=None):
    _check(load_library().rk_optimizer_step(
        _stream(), OPT_KIND[kind], _ptr(hyper), n, _ptr(params), _ptr(state), _ptr(grad),
        _ptr(lo), _ptr(hi), _ptr(zero_flag)), "rk_optimizer_step")


def reduce_optimizer_step(gradp, rows: int, n: int, grad, kind: str, hyper, params, state, lo, hi,
                          zero_flag, counters):
    """rk_reduce_optimizer_step: partial-row reduction and optimiser step in one launch."""
    _check(load_library().rk_reduce_optimizer_step(
        _stream(), _ptr(gradp), rows, n, _ptr(grad), OPT_KIND[kind], _ptr(hyper), _ptr(params),
        _ptr(state), _ptr(lo), _ptr(hi), _ptr(zero_flag), _ptr(counters)), "rk_reduce_optimizer_step")


def optimizer_chain_step(kind: str, hyper: torch.Tensor, n: int, params, state, grad, lo, hi, zero_flag,
                         cfg: dict, count, plateau, ema, loss, seed: int, scratch):
    """rk_optimizer_chain_step: the optimiser with the add_noise / reduce_on_plateau / ema links of the
    optax chain.  ``cfg``: noise = (eta, gamma) or None, ema_decay or None, plateau = the seven
    reduce_on_plateau arguments or None."""
    noise, pk, d = cfg.get("noise"), cfg.get("plateau"), cfg.get("ema_decay")
    vals = [noise[0] if noise else 0.0, noise[1] if noise else 0.0, d if d is not None else 0.0]
    vals += [float(pk[k]) if pk else 0.0 for k in ("factor", "patience", "rtol", "atol", "cooldown",
                                                    "accumulation_size", "min_scale")]
    vals += [1.0 if noise else 0.0, 1.0 if pk else 0.0, 1.0 if d is not None else 0.0]
    arr = (ctypes.c_float * 13)(*vals)
    _check(load_library().rk_optimizer_chain_step(
        _stream(), OPT_KIND[kind], _ptr(hyper), n, _ptr(params), _ptr(state), _ptr(grad), _ptr(lo),
        _ptr(hi), _ptr(zero_flag), ctypes.cast(arr, ctypes.c_void_p), _ptr(count), _ptr(plateau),
        _ptr(ema), _ptr(loss), ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), _ptr(scratch)),
        "rk_optimizer_chain_step")


def mean_loss(returns: torch.Tensor, B: int, loss=None, gret=None):
    _check(load_library().rk_mean_loss(_stream(), _ptr(returns), B, _ptr(loss), _ptr(gret)),
           "rk_mean_loss")


def project_slp(method: str, T: int, A: int, params: torch.Tensor, segs, allowed: int,
                max_iter: int, wrap_sigmoid: bool, min_prob: float, converged=None):
    """segs: [(offset, length, noop, sigmoid_weight), ...] of the boolean action fluents."""
    n = len(segs)
    I32, F32 = ctypes.c_int32 * max(n, 1), ctypes.c_float * max(n, 1)
    off, ln = I32(*[int(s[0]) for s in segs]), I32(*[int(s[1]) for s in segs])
    noop, w = I32(*[int(bool(s[2])) for s in segs]), F32(*[float(s[3]) for s in segs])
    _check(load_library().rk_project_slp(
        _stream(), {"sogbofa": 0, "sorting": 1}[method], T, A, _ptr(params), n,
        ctypes.cast(off, ctypes.c_void_p), ctypes.cast(ln, ctypes.c_void_p),
        ctypes.cast(noop, ctypes.c_void_p), ctypes.cast(w, ctypes.c_void_p), int(allowed),
        int(max_iter), int(bool(wrap_sigmoid)), float(min_prob), _ptr(converged)),
        "rk_project_slp")


GEMM_A_TRANS, GEMM_ACCUMULATE = 1, 2
EPI_NONE, EPI_BIAS, EPI_BIAS_TANH, EPI_DTANH = 0 << 4, 1 << 4, 2 << 4, 3 << 4


def sgemm(flags: int, M: int, N: int, K: int, A, lda: int, Bm, ldb: int, C, ldc: int, bias=None,
          aux=None, ldaux: int = 0, workspace=None, split_k: int = 1, act=None):
    _check(load_library().rk_sgemm(_stream(), flags | (_act_id(act) << 8), M, N, K, _ptr(A), lda, _ptr(Bm), ldb,
                                   _ptr(C), ldc, _ptr(bias), _ptr(aux), ldaux, _ptr(workspace),
                                   split_k), "rk_sgemm")


def _segs(segs):
    n = len(segs)
    I32 = ctypes.c_int32 * max(n, 1)
    off, ln = I32(*[int(s[0]) for s in segs]), I32(*[int(s[1]) for s in segs])
    return n, off, ln


def _kinds(segs, kinds):
    I32 = ctypes.c_int32 * max(len(segs), 1)
    return I32(*[int(k) for k in (kinds if kinds is not None else [0] * len(segs))])


def drp_actions_forward(B, A, stochastic, segs, heads, noise, train, lo, hi, ls_lo, ls_hi, actions,
                        entropy=None, kinds=None, bool_weight: float = 1.0, typed: bool = False,
                        wrap: bool = False):
    """segs: [(word offset, length), ...] of the action fluents (fluent-major layout); kinds:
    0 real / 1 bool / 2 int per fluent; typed: write bool / int actions as int32 words (the exact
    test program's input) instead of floats."""
    n, off, ln = _segs(segs)
    kd = _kinds(segs, kinds)
    _check(load_library().rk_drp_actions_forward(
        _stream(), B, A, int(stochastic), n, ctypes.cast(off, ctypes.c_void_p),
        ctypes.cast(ln, ctypes.c_void_p), ctypes.cast(kd, ctypes.c_void_p), _ptr(heads),
        _ptr(noise), float(train), _ptr(lo), _ptr(hi), float(ls_lo), float(ls_hi),
        float(bool_weight), int(bool(typed)) | (2 if wrap else 0), _ptr(actions), _ptr(entropy)),
        "rk_drp_actions_forward")


def drp_actions_backward(B, A, stochastic, segs, heads, noise, train, lo, hi, ls_lo, ls_hi, gact,
                         gheads, kinds=None, bool_weight: float = 1.0, ent_coeff=None,
                         inv_batch: float = 0.0, wrap: bool = False,
                         sigma_entropy_grad: bool = False):
    n, off, ln = _segs(segs)
    kd = _kinds(segs, kinds)
    _check(load_library().rk_drp_actions_backward(
        _stream(), B, A, int(stochastic), n, ctypes.cast(off, ctypes.c_void_p),
        ctypes.cast(ln, ctypes.c_void_p), ctypes.cast(kd, ctypes.c_void_p), _ptr(heads),
        _ptr(noise), float(train), _ptr(lo), _ptr(hi), float(ls_lo), float(ls_hi),
        float(bool_weight), _ptr(ent_coeff), float(inv_batch),
        int(bool(wrap)) | (2 if sigma_entropy_grad else 0), _ptr(gact),
        _ptr(gheads)),
        "rk_drp_actions_backward")


def state_scale(B, D, segs, lo, hi, x, y, backward: bool = False):
    """forward: y = (x - lo) / (hi - lo) per state word; backward: y += x / (hi - lo)."""
    n, off, ln = _segs(segs)
    lib = load_library()
    fn = lib.rk_state_scale_backward if backward else lib.rk_state_scale_forward
    _check(fn(_stream(), B, D, n, ctypes.cast(off, ctypes.c_void_p), ctypes.cast(ln, ctypes.c_void_p),
              _ptr(lo), _ptr(hi), _ptr(x), _ptr(y)), "rk_state_scale")


ACTIVATIONS = {"tanh": 0, "relu": 1, "sigmoid": 2, "softplus": 3, "elu": 4, "leaky_relu": 5,
               "gelu": 6, "silu": 7, "swish": 7}
# activations whose derivative needs the pre-activation: the layer-wise path keeps z beside act(z)
Z_ACTIVATIONS = ("gelu", "silu", "swish")
ACT_IDENTITY = 8


def act_apply(act, n: int, z, y):
    """y = act(z) element-wise (the second half of a gelu / silu layer)."""
    _check(load_library().rk_act_apply(_stream(), _act_id(act), ctypes.c_size_t(int(n)), _ptr(z), _ptr(y)),
           "rk_act_apply")


_ACT = 0


def drp_set_activation(name: str) -> None:
    """Default hidden-layer activation of the wrappers below (a host-side convenience of this binding: the
    library itself takes the activation id as an argument of every launch, include/rk.h)."""
    global _ACT
    _ACT = ACTIVATIONS[name]


def _act_id(act) -> int:
    return _ACT if act is None else (ACTIVATIONS[act] if isinstance(act, str) else int(act))


def film_chunks(B: int) -> int:
    return int(load_library().rk_film_chunks(int(B)))


def film_forward(B, H, z, gamma, beta, out):
    _check(load_library().rk_film_forward(_stream(), B, H, _ptr(z), _ptr(gamma), _ptr(beta),
                                          _ptr(out)), "rk_film_forward")


def film_backward(B, H, g, z, gamma, gz, grad_gamma_beta, workspace, act=None):
    _check(load_library().rk_film_backward(_stream(), B, H, _ptr(g), _ptr(z), _ptr(gamma), _ptr(gz),
                                           _ptr(grad_gamma_beta), _ptr(workspace), _act_id(act)),
           "rk_film_backward")


TOPK_MAX_POOLED, TOPK_MAX_ALLOWED = 64, 8      # RK_TOPK_MAX_N / RK_TOPK_MAX_K of csrc/rk_drp.cu


def drp_topk_limits():
    a, b = ctypes.c_int(0), ctypes.c_int(0)
    _check(load_library().rk_drp_topk_limits(ctypes.byref(a), ctypes.byref(b)), "rk_drp_topk_limits")
    return a.value, b.value


def drp_topk_forward(B, A, stochastic, segs, kinds, heads, noise, train, allowed, weight, actions,
                     entropy=None, typed: bool = False):
    """Pooled boolean heads (kinds 3 / 4) under max-nondef-actions = allowed; see include/rk.h."""
    n, off, ln = _segs(segs)
    kd = _kinds(segs, kinds)
    _check(load_library().rk_drp_topk_forward(
        _stream(), B, A, int(stochastic), n, ctypes.cast(off, ctypes.c_void_p),
        ctypes.cast(ln, ctypes.c_void_p), ctypes.cast(kd, ctypes.c_void_p), _ptr(heads),
        _ptr(noise), int(bool(train)), int(allowed), float(weight), int(bool(typed)),
        _ptr(actions), _ptr(entropy)), "rk_drp_topk_forward")


def drp_topk_backward(B, A, stochastic, segs, kinds, heads, noise, allowed, weight, gact, gheads,
                      ent_coeff=None, inv_batch: float = 0.0):
    n, off, ln = _segs(segs)
    kd = _kinds(segs, kinds)
    _check(load_library().rk_drp_topk_backward(
        _stream(), B, A, int(stochastic), n, ctypes.cast(off, ctypes.c_void_p),
        ctypes.cast(ln, ctypes.c_void_p), ctypes.cast(kd, ctypes.c_void_p), _ptr(heads),
        _ptr(noise), int(allowed), float(weight), _ptr(ent_coeff), float(inv_batch), _ptr(gact),
        _ptr(gheads)), "rk_drp_topk_backward")


def tf32_split(x: torch.Tensor, hi: torch.Tensor, lo: torch.Tensor, n: Optional[int] = None):
    _check(load_library().rk_tf32_split(_stream(), _ptr(x), _ptr(hi), _ptr(lo),
                                        int(x.numel() if n is None else n)), "rk_tf32_split")


def tcgemm(epi: int, M: int, N: int, K: int, A, lda: int, Bt_hi, Bt_lo, ldb: int, C,
           ldc: int, bias=None, aux=None, ldaux: int = 0, act=None):
    """Tensor-core (tcgen05, 3xTF32) C = epi(A * Bt^T); epi as EPI_* >> 4."""
    _check(load_library().rk_tcgemm(_stream(), epi | (_act_id(act) << 8), M, N, K, _ptr(A), lda,
                                    _ptr(Bt_hi), _ptr(Bt_lo), ldb, _ptr(C), ldc, _ptr(bias),
                                    _ptr(aux), ldaux), "rk_tcgemm")


def drp_mlp2_image_floats(D, h0, h1, NH, backward: bool) -> int:
    n = int(load_library().rk_drp_mlp2_image_floats(D, h0, h1, NH, 1 if backward else 0))
    if n < 0:
        raise RkError(f"rk_drp_mlp2_image_floats: unsupported shape {(D, h0, h1, NH)}")
    return n


def drp_mlp2_prepare(D, h0, h1, NH, W0, Wh, fwd_img, bwd_img):
    """Operand images of the fused policy kernels (once per optimiser step)."""
    _check(load_library().rk_drp_mlp2_prepare(_stream(), D, h0, h1, NH, _ptr(W0), _ptr(Wh),
                                              _ptr(fwd_img), _ptr(bwd_img)), "rk_drp_mlp2_prepare")


def drp_mlp2_forward(act, B, D, h0, h1, NH, segs, x, fwd_img, b0, W1T_hi, W1T_lo, b1, bh, H0, H1, heads):
    """Fused forward of a 2-hidden-layer policy network (one launch per decision epoch);
    ``act`` is the activation's name."""
    n, off, ln = _segs(segs)
    _check(load_library().rk_drp_mlp2_forward(
        _stream(), ACTIVATIONS[act], B, D, h0, h1, NH, n, ctypes.cast(off, ctypes.c_void_p),
        ctypes.cast(ln, ctypes.c_void_p), _ptr(x), _ptr(fwd_img), _ptr(b0), _ptr(W1T_hi), _ptr(W1T_lo),
        _ptr(b1), _ptr(bh), _ptr(H0), _ptr(H1), _ptr(heads)), "rk_drp_mlp2_forward")


def drp_mlp2_backward(act, B, D, h0, h1, NH, segs, gheads, H0, H1, bwd_img, W1_hi, W1_lo, gz0, gz1, gs):
    """Fused input-side adjoint of a 2-hidden-layer policy network: pre-activation cotangents gz1, gz0
    (written once) and the input cotangent added to the fluent-major state cotangent ``gs``."""
    n, off, ln = _segs(segs)
    _check(load_library().rk_drp_mlp2_backward(
        _stream(), ACTIVATIONS[act], B, D, h0, h1, NH, n, ctypes.cast(off, ctypes.c_void_p),
        ctypes.cast(ln, ctypes.c_void_p), _ptr(gheads), _ptr(H0), _ptr(H1), _ptr(bwd_img), _ptr(W1_hi),
        _ptr(W1_lo), _ptr(gz0), _ptr(gz1), _ptr(gs)), "rk_drp_mlp2_backward")


NOISE_KINDS = {"uniform": 0, "normal": 1, "exponential": 2, "gumbel": 3, "logistic": 4, "laplace": 5,
               "cauchy": 6, "gamma": 7}


def draw_noise(T: int, B: int, sites, seed: int, draws: Optional[torch.Tensor], out: torch.Tensor,
               conc: Optional[torch.Tensor] = None):
    """out[T][N * B] <- fresh noise of the draw sites ``sites`` = [(kind name, first word, words)], one
    launch (Philox4x32-10 + the site's law); ``draws``: int64 device word with the draw number; ``conc``:
    [N] device floats, the Gamma concentration of every noise word (read at the gamma sites only)."""
    n = len(sites)
    kinds = (ctypes.c_int32 * n)(*[NOISE_KINDS[k] for k, _, _ in sites])
    offs = (ctypes.c_int32 * (n + 1))(*([o for _, o, _ in sites] + [sites[-1][1] + sites[-1][2]]))
    _check(load_library().rk_draw_noise(
        _stream(), T, B, n, ctypes.cast(kinds, ctypes.c_void_p), ctypes.cast(offs, ctypes.c_void_p),
        ctypes.c_ulonglong(seed & 0xFFFFFFFFFFFFFFFF), _ptr(draws), _ptr(out), _ptr(conc)), "rk_draw_noise")


def drp_returns_loss(T, B, reward, disc, entropy, ent_coeff, returns, ent_scratch, loss):
    """returns[b] = sum_t disc_t reward[t][b]; loss = -mean(returns) - ent_coeff * mean_b sum_t entropy[t][b]."""
    _check(load_library().rk_drp_returns_loss(
        _stream(), T, B, _ptr(reward), _ptr(disc), _ptr(entropy), _ptr(ent_coeff), _ptr(returns),
        _ptr(ent_scratch), _ptr(loss)), "rk_drp_returns_loss")


UTILITY_KINDS = {"mean_var": 1, "mean_std": 2, "mean_semivar": 3, "mean_semidev": 4, "sharpe": 5,
                 "entropic": 6, "exponential": 6, "symlog_mean": 7}
# the order-statistic utilities: one device sort of the return vector (csrc/rk_utility.cu)
UTILITY_SORTED_KINDS = {"var": 8, "cvar": 9, "cvar_ste": 10, "expectile": 11, "gini": 12}


def utility_loss(kind: str, B: int, returns, p0: float, p1: float, scale: float, loss, gret,
                 b0: int = 0, n_local: Optional[int] = None):
    """loss = -U(returns), gret = -scale * dU/dreturns for the closed-form utilities (include/rk.h);
    ``returns`` is the global vector [B], gret covers the rollouts [b0, b0 + n_local)."""
    n_local = B if n_local is None else n_local
    _check(load_library().rk_utility_loss_shard(
        _stream(), UTILITY_KINDS[kind], B, int(b0), int(n_local), _ptr(returns), float(p0), float(p1),
        float(scale), _ptr(loss), _ptr(gret)), "rk_utility_loss_shard")


def utility_sorted_work(B: int, device) -> torch.Tensor:
    n = int(load_library().rk_utility_sorted_work_bytes(int(B)))
    return torch.empty(n, dtype=torch.uint8, device=device)


def utility_sorted_loss(kind: str, B: int, returns, p0: float, max_iter: int, scale: float, loss, gret,
                        work: torch.Tensor, b0: int = 0, n_local: Optional[int] = None):
    """var / cvar / cvar_ste / expectile / gini and their cotangents (include/rk.h)."""
    n_local = B if n_local is None else n_local
    _check(load_library().rk_utility_sorted_loss(
        _stream(), UTILITY_SORTED_KINDS[kind], B, int(b0), int(n_local), _ptr(returns), float(p0),
        int(max_iter), float(scale), _ptr(loss), _ptr(gret), _ptr(work), work.numel()),
        "rk_utility_sorted_loss")


def transpose_split(r, c, W, WT, WT_hi, WT_lo):
    """W [r][c] -> W^T (and / or its TF32 hi / lo split)."""
    _check(load_library().rk_transpose_split(_stream(), r, c, _ptr(W), _ptr(WT), _ptr(WT_hi),
                                             _ptr(WT_lo)), "rk_transpose_split")


def noise_advance(draws: torch.Tensor):
    _check(load_library().rk_noise_advance(_stream(), _ptr(draws)), "rk_noise_advance")


def drp_mlp2_set_timing(buf: Optional[torch.Tensor]) -> None:
    """Profiling aid: int64 device tensor of >= grid * 8 words receiving per-CTA phase stamps, or None."""
    _check(load_library().rk_drp_mlp2_set_timing(_ptr(buf)), "rk_drp_mlp2_set_timing")


def drp_wgrad_ctas() -> int:
    return int(load_library().rk_drp_wgrad_ctas())


def drp_wgrad_hidden(accumulate, M, N, K, A, Bm, dW, db, workspace):
    """dW [M][N] (+)= A[K][M]^T B[K][N]; db [N] (+)= column sums of B (or None)."""
    _check(load_library().rk_drp_wgrad_hidden(_stream(), 1 if accumulate else 0, M, N, K, _ptr(A),
                                              _ptr(Bm), _ptr(dW), _ptr(db), _ptr(workspace)),
           "rk_drp_wgrad_hidden")


def drp_wgrad_heads(accumulate, M, NH, K, H, G, dWh, dbh, workspace):
    """dWh [M][NH] (+)= H^T G; dbh [NH] (+)= column sums of G (or None)."""
    _check(load_library().rk_drp_wgrad_heads(_stream(), 1 if accumulate else 0, M, NH, K, _ptr(H),
                                             _ptr(G), _ptr(dWh), _ptr(dbh), _ptr(workspace)),
           "rk_drp_wgrad_heads")


def drp_wgrad_input(accumulate, M, D, T, Bt, S_words, segs, gz0, x_tape, dW0, db0, workspace):
    n, off, ln = _segs(segs)
    _check(load_library().rk_drp_wgrad_input(
        _stream(), 1 if accumulate else 0, M, D, T, Bt, S_words, n, ctypes.cast(off, ctypes.c_void_p),
        ctypes.cast(ln, ctypes.c_void_p), _ptr(gz0), _ptr(x_tape), _ptr(dW0), _ptr(db0),
        _ptr(workspace)), "rk_drp_wgrad_input")


def drp_layer0(B, D, H, segs, x, W0, b0, out, xp=None, act=None):
    n, off, ln = _segs(segs)
    _check(load_library().rk_drp_layer0(
        _stream(), B, D, H, n, ctypes.cast(off, ctypes.c_void_p), ctypes.cast(ln, ctypes.c_void_p),
        _ptr(x), _ptr(W0), _ptr(b0), _ptr(out), _ptr(xp), _act_id(act)), "rk_drp_layer0")


def drp_head_backward_ctas(B: int) -> int:
    return int(load_library().rk_drp_head_backward_ctas(int(B)))


def drp_head_backward(B, hl, NH, Hl, G, Wh, gz, grad_tail, workspace, act=None):
    _check(load_library().rk_drp_head_backward(
        _stream(), B, hl, NH, _ptr(Hl), _ptr(G), _ptr(Wh), _ptr(gz), _ptr(grad_tail),
        _ptr(workspace), _act_id(act)), "rk_drp_head_backward")


def drp_layer0_backward(B, D, H, segs, x, gz0, W0, gs, grad_W0b0, workspace):
    n, off, ln = _segs(segs)
    _check(load_library().rk_drp_layer0_backward(
        _stream(), B, D, H, n, ctypes.cast(off, ctypes.c_void_p), ctypes.cast(ln, ctypes.c_void_p),
        _ptr(x), _ptr(gz0), _ptr(W0), _ptr(gs), _ptr(grad_W0b0), _ptr(workspace)),
        "rk_drp_layer0_backward")


def peer_allreduce(world: int, rank: int, peer_bufs, peer_flags, n: int, n_max: int, vec, epoch,
                   status=None):
    """peer_bufs / peer_flags: ctypes arrays of uint64 device addresses (one per rank)."""
    _check(load_library().rk_peer_allreduce(
        _stream(), world, rank, ctypes.cast(peer_bufs, ctypes.c_void_p),
        ctypes.cast(peer_flags, ctypes.c_void_p), n, n_max, _ptr(vec), _ptr(epoch), _ptr(status)),
        "rk_peer_allreduce")


def peer_allreduce_gather(world: int, rank: int, peer_bufs, peer_flags, n: int, n_max: int, vec, epoch,
                          status, extras, dst):
    """The exchange with a gathered tail: ``extras`` = short device vectors appended to the message,
    their sums land in ``dst`` (contiguous)."""
    k = len(extras)
    ptrs = (ctypes.c_void_p * max(k, 1))(*[_ptr(t) for t in extras])
    lens = (ctypes.c_int32 * max(k, 1))(*[int(t.numel()) for t in extras])
    _check(load_library().rk_peer_allreduce_gather(
        _stream(), world, rank, ctypes.cast(peer_bufs, ctypes.c_void_p),
        ctypes.cast(peer_flags, ctypes.c_void_p), n, n_max, _ptr(vec), _ptr(epoch), _ptr(status), k,
        ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(lens, ctypes.c_void_p), _ptr(dst)),
        "rk_peer_allreduce_gather")


def state_layernorm_ctas(B: int) -> int:
    return int(load_library().rk_state_layernorm_ctas(int(B)))


def state_layernorm_forward(B, D, segs, norm, eps, x, scale, bias, y):
    n, off, ln = _segs(segs)
    kd = _kinds(segs, [int(k) for k in norm])
    _check(load_library().rk_state_layernorm_forward(
        _stream(), B, D, n, ctypes.cast(off, ctypes.c_void_p), ctypes.cast(ln, ctypes.c_void_p),
        ctypes.cast(kd, ctypes.c_void_p), float(eps), _ptr(x), _ptr(scale), _ptr(bias), _ptr(y)),
        "rk_state_layernorm_forward")


def state_layernorm_backward(B, D, segs, norm, eps, x, scale, gy, gx, grad_scale_bias, workspace):
    n, off, ln = _segs(segs)
    kd = _kinds(segs, [int(k) for k in norm])
    _check(load_library().rk_state_layernorm_backward(
        _stream(), B, D, n, ctypes.cast(off, ctypes.c_void_p), ctypes.cast(ln, ctypes.c_void_p),
        ctypes.cast(kd, ctypes.c_void_p), float(eps), _ptr(x), _ptr(scale), _ptr(gy), _ptr(gx),
        _ptr(grad_scale_bias), _ptr(workspace)), "rk_state_layernorm_backward")


def colsum(B, N, G, ldg, y, accumulate, workspace, chunks):
    _check(load_library().rk_colsum(_stream(), B, N, _ptr(G), ldg, _ptr(y), int(bool(accumulate)),
                                    _ptr(workspace), int(chunks)), "rk_colsum")


def rowdot(accumulate, B, N, K, A, lda, W, ldw, out, ldo, bias=None):
    _check(load_library().rk_rowdot(_stream(), int(bool(accumulate)), B, N, K, _ptr(A), lda,
                                    _ptr(W), ldw, _ptr(out), ldo, _ptr(bias)), "rk_rowdot")


def tcgemm_tn(accumulate, M, N, K, A, lda, Bm, ldb, C, workspace, max_splits):
    """Tensor-core (tcgen05, 3xTF32) C (+)= A^T * B, both operands [K][.] row-major."""
    _check(load_library().rk_tcgemm_tn(_stream(), int(bool(accumulate)), M, N, K, _ptr(A), lda,
                                       _ptr(Bm), ldb, _ptr(C), _ptr(workspace), int(max_splits)),
           "rk_tcgemm_tn")


def tile_state(B, segs, init, s0):
    """segs: [(word offset, length), ...] of the state fluents; init: [S] device words."""
    n, off, ln = _segs(segs)
    _check(load_library().rk_tile_state(_stream(), B, n, ctypes.cast(off, ctypes.c_void_p),
                                        ctypes.cast(ln, ctypes.c_void_p), _ptr(init), _ptr(s0)),
           "rk_tile_state")

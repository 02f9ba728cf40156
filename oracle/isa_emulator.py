"""TEST INFRASTRUCTURE: a numpy emulator of the rollout-program ISA (core/isa.py).

It decodes the binary blob exactly as csrc/rk_interp.cu does and executes it on the CPU,
vectorised over warps, so that lowering, differentiation, register allocation and encoding
can be checked against the oracle without a GPU (``pytest -m "not gpu"``).  It is NOT a
product path: nothing in pyrddlgym_jax_b200/ imports it.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from pyrddlgym_jax_b200.core import isa

H = 32


def _stable_tanh(x):
    ax = np.abs(x)
    with np.errstate(over="ignore", invalid="ignore"):
        em = np.expm1(2 * ax)
        small = np.where(ax < 20, em / (em + 2), 1 - 2 * np.exp(-2 * ax))
    return (np.sign(x) * small).astype(np.float32)


def _pymod(x, y):
    return np.mod(x, y)


class Emulator:
    def __init__(self, blob: bytes):
        words = np.frombuffer(blob, dtype=np.uint32)
        self.hdr = words[:H].copy()
        assert self.hdr[0] == isa.MAGIC
        self.n_pro, self.n_step, self.n_epi = (int(x) for x in self.hdr[2:5])
        self.cw = int(self.hdr[5])
        self.mparam_off, self.n_mp = int(self.hdr[6]), int(self.hdr[7])
        self.ww, self.rpw, self.L = int(self.hdr[8]), int(self.hdr[9]), int(self.hdr[10])
        self.S, self.A, self.N = int(self.hdr[11]), int(self.hdr[12]), int(self.hdr[13])
        self.backward = bool(self.hdr[14] & 2)
        self.static_err, self.err_base = int(self.hdr[15]), int(self.hdr[16])
        self.traj_w = int(self.hdr[17])
        self.tape_w = int(self.hdr[18]) or int(self.hdr[11])
        self.C0 = words[H:H + self.cw].copy()
        ins = np.frombuffer(blob, dtype=np.uint16, offset=(H + self.cw) * 4)
        self.ins = ins.reshape(-1, isa.INS_WORDS).astype(np.int64)

    def num_warps(self, B):
        return -(-B // self.rpw)

    def run(self, B: int, T: int, bufs: Dict[str, Optional[np.ndarray]],
            mparams: Optional[np.ndarray] = None):
        """bufs: name -> flat numpy array (float32/int32/uint32 views are all fine)."""
        rpw, nW = self.rpw, self.num_warps(B)
        C = self.C0.copy()
        if mparams is not None:
            C[self.mparam_off:self.mparam_off + self.n_mp] = \
                np.asarray(mparams, np.float32).view(np.uint32)
        R = np.zeros((nW, self.ww), dtype=np.uint32)
        Wd = {"S0": self.S, "TAPE": self.tape_w, "NOISE": self.N, "PARAMS": self.A, "ACT": self.A,
              "REWARD": 1, "RETURNS": 1, "ERR": 1, "DISC": 1, "FINAL": self.S,
              "TRAJ": self.traj_w, "GRET": 1, "GRADP": self.A, "GACT": self.A, "GS0": self.S,
              "GSIN": self.S}
        G = {}
        for name in isa.BUFFERS:
            a = bufs.get(name)
            G[isa.BUF[name]] = None if a is None else a.reshape(-1).view(np.uint32)
        b0 = np.arange(nW) * rpw
        nvalid = np.minimum(rpw, B - b0)
        valid_q = np.arange(rpw)[None, :] < nvalid[:, None]          # [nW, rpw]
        ctx = dict(B=B, T=T, C=C, R=R, G=G, Wd={isa.BUF[k]: v for k, v in Wd.items()}, b0=b0,
                   nvalid=nvalid, valid_q=valid_q)
        self._run_list(0, self.n_pro, 0, ctx)
        ts = range(T - 1, -1, -1) if self.backward else range(T)
        for t in ts:
            self._run_list(self.n_pro, self.n_step, t, ctx)
        self._run_list(self.n_pro + self.n_step, self.n_epi, 0, ctx)

    # ---------------------------------------------------------------------------------
    def _addr(self, base, sq, mp, n, C, uniform_dst=False):
        """Word addresses [rpw, n] (or [1, n] when sq == 0) and whether const space."""
        const = bool(base & isa.CONST_BIT)
        b = base & 0x7FFF if const else base
        if mp == 0:
            off = np.arange(n)
        elif mp == 1:
            off = np.zeros(n, dtype=np.int64)
        else:
            off = C[mp:mp + n].astype(np.int64)
        q = np.arange(self.rpw)[:, None] if sq else np.zeros((1, 1), dtype=np.int64)
        return const, b + q * sq + off[None, :]

    def _load(self, w, k, n, ctx):
        base, sq, mp = w[5 + 3 * k: 8 + 3 * k]
        const, ad = self._addr(int(base), int(sq), int(mp), n, ctx["C"])
        if const:
            v = ctx["C"][ad]                     # [1 or rpw, n]
            return np.broadcast_to(v[None], (ctx["R"].shape[0],) + v.shape)
        return ctx["R"][:, ad]                   # [nW, 1 or rpw, n]

    def _run_list(self, start, count, t, ctx):
        R, C, G, B, T = ctx["R"], ctx["C"], ctx["G"], ctx["B"], ctx["T"]
        rpw = self.rpw
        f32 = lambda u: u.view(np.float32)          # noqa: E731
        for w in self.ins[start:start + count]:
            op = isa.OPS[int(w[0])]
            flags, n, dst, dst_sq = int(w[1]), int(w[2]), int(w[3]), int(w[4])
            aux0, aux1 = int(w[14]), int(w[15])
            uni = bool(flags & isa.F_UNIFORM)

            def store(val):
                val = np.ascontiguousarray(val)
                u = val.view(np.uint32) if val.dtype != np.uint32 else val
                if dst_sq == 0:
                    R[:, dst:dst + n] = u.reshape(u.shape[0], -1, n)[:, 0, :]
                else:
                    ad = dst + np.arange(rpw)[:, None] * dst_sq + np.arange(n)[None, :]
                    R[:, ad] = np.broadcast_to(u, (R.shape[0], rpw, n))

            if op == "NOP":
                continue
            if op in ("LDB", "STB", "LDU", "STG", "STERR"):
                buf = aux0
                off = aux1 | (int(w[11]) << 16)
                g = G.get(buf)
                tt = t if flags & isa.F_TSTEP else 0
                Wb = ctx["Wd"][buf]
                if op == "STERR":
                    e = R[:, self.err_base:self.err_base + rpw] | np.uint32(self.static_err)
                    if g is not None:
                        idx = (tt * B + ctx["b0"][:, None] + np.arange(rpw)[None, :])
                        g[idx[ctx["valid_q"]]] = e[ctx["valid_q"]]
                    R[:, self.err_base:self.err_base + rpw] = 0
                    continue
                if g is None:
                    continue
                if op == "LDU":
                    R[:, dst:dst + n] = g[tt * Wb + off: tt * Wb + off + n][None, :]
                    continue
                if op == "STG":
                    v = self._load(w, 0, n, ctx)[:, 0, :]
                    row = (np.arange(R.shape[0]) * T + t) * Wb + off
                    g[row[:, None] + np.arange(n)[None, :]] = v
                    continue
                gidx = (tt * Wb + off) * B + (ctx["b0"][:, None, None] + np.arange(rpw)[None, :, None]) * n \
                    + np.arange(n)[None, None, :]
                mask = np.broadcast_to(ctx["valid_q"][:, :, None], gidx.shape)
                if op == "LDB":
                    ad = dst + np.arange(rpw)[:, None] * n + np.arange(n)[None, :]
                    cur = R[:, ad]
                    cur[mask] = g[gidx[mask]]
                    R[:, ad] = cur
                else:
                    v = np.broadcast_to(self._load(w, 0, n, ctx), gidx.shape)
                    g[gidx[mask]] = v[mask]
                continue
            if op == "ERRCHK":
                v = np.broadcast_to(self._load(w, 0, n, ctx), (R.shape[0], rpw, n))
                bad = (v != 0).any(axis=2) & ctx["valid_q"]
                R[:, self.err_base:self.err_base + rpw] |= (bad.astype(np.uint32) << np.uint32(aux0))
                continue
            if op == "QSUM":
                a, a_sq = int(w[5]), int(w[6])
                ad = a + np.arange(rpw)[:, None] * a_sq + np.arange(n)[None, :]
                v = f32(R[:, ad]).copy()
                v[~ctx["valid_q"]] = 0
                acc = np.zeros((R.shape[0], n), dtype=np.float32)
                for qq in range(rpw):
                    acc = acc + v[:, qq, :]
                R[:, dst:dst + n] = acc.view(np.uint32)
                continue
            if op.startswith("R") and op not in ("RELU", "RCP", "ROUND"):
                base, sq = int(w[5]), int(w[6])
                nseg = n
                if op == "RPRODX":      # n counts source elements; the CSR has one row per segment
                    nseg = int(C[int(w[8]):int(w[8]) + n].astype(np.int64).max()) + 1 if n else 0
                ptr = C[aux0:aux0 + nseg + 1].astype(np.int64)
                idx = C[aux1:aux1 + int(ptr[-1])].astype(np.int64)
                qs = np.arange(rpw)[:, None] if sq else np.zeros((1, 1), dtype=np.int64)
                const = bool(base & isa.CONST_BIT)
                src = C if const else None
                out = None
                is_int = op in ("RISUM", "RIPROD", "RIMIN", "RIMAX")
                if op == "RPRODX":
                    seg = C[int(w[8]):int(w[8]) + n].astype(np.int64)
                res = []
                for e in range(n):
                    if op == "RPRODX":
                        s = seg[e]
                        members = idx[ptr[s]:ptr[s + 1]]
                        members = members[members != e]
                    else:
                        members = idx[ptr[e]:ptr[e + 1]]
                    ad = (base & 0x7FFF) + qs * sq + members[None, :]
                    vals = (np.broadcast_to(C[ad][None], (R.shape[0],) + ad.shape) if const
                            else R[:, ad])
                    vals = vals.view(np.int32) if (is_int) else vals.view(np.float32)
                    if op in ("RSUM", "RISUM"):
                        acc = np.zeros(vals.shape[:2], dtype=vals.dtype)
                        for j in range(vals.shape[2]):
                            acc = acc + vals[:, :, j]
                    elif op in ("RPROD", "RIPROD", "RPRODX"):
                        acc = np.ones(vals.shape[:2], dtype=vals.dtype)
                        for j in range(vals.shape[2]):
                            acc = acc * vals[:, :, j]
                    elif op in ("RMIN", "RIMIN"):
                        acc = vals.min(axis=2) if vals.shape[2] else np.full(
                            vals.shape[:2], np.inf if not is_int else 2**31 - 1, dtype=vals.dtype)
                    elif op in ("RMAX", "RIMAX"):
                        acc = vals.max(axis=2) if vals.shape[2] else np.full(
                            vals.shape[:2], -np.inf if not is_int else -2**31, dtype=vals.dtype)
                    elif op == "RARGMAX":
                        acc = np.argmax(vals, axis=2).astype(np.int32)
                    elif op == "RARGMIN":
                        acc = np.argmin(vals, axis=2).astype(np.int32)
                    else:
                        raise NotImplementedError(op)
                    res.append(acc)
                out = np.stack(res, axis=2)
                store(out)
                continue
            # ---- elementwise
            x = self._load(w, 0, n, ctx)
            y = self._load(w, 1, n, ctx)
            z = self._load(w, 2, n, ctx)
            xf, yf, zf = f32(np.ascontiguousarray(x)), f32(np.ascontiguousarray(y)), \
                f32(np.ascontiguousarray(z))
            xi, yi = np.ascontiguousarray(x).view(np.int32), np.ascontiguousarray(y).view(np.int32)
            one = np.float32(1)
            with np.errstate(all="ignore"):
                if op == "MOV": r = x
                elif op == "NEG": r = -xf
                elif op == "ABS": r = np.abs(xf)
                elif op == "SGN": r = np.sign(xf)
                elif op == "FLOOR": r = np.floor(xf)
                elif op == "CEIL": r = np.ceil(xf)
                elif op == "ROUND": r = np.round(xf)
                elif op == "SQRT": r = np.sqrt(xf)
                elif op == "EXP": r = np.exp(xf)
                elif op == "LOG": r = np.log(xf)
                elif op == "SIN": r = np.sin(xf)
                elif op == "COS": r = np.cos(xf)
                elif op == "TAN": r = np.tan(xf)
                elif op == "ASIN": r = np.arcsin(xf)
                elif op == "ACOS": r = np.arccos(xf)
                elif op == "ATAN": r = np.arctan(xf)
                elif op == "SINH": r = np.sinh(xf)
                elif op == "COSH": r = np.cosh(xf)
                elif op == "TANH": r = np.tanh(xf)
                elif op == "LGAMMA":
                    from scipy.special import gammaln
                    r = gammaln(xf).astype(np.float32)
                elif op == "SIGMOID": r = one / (one + np.exp(-xf))
                elif op == "STANH": r = _stable_tanh(xf)
                elif op == "RELU": r = np.maximum(xf, np.float32(0))
                elif op == "RCP": r = one / xf
                elif op == "SQUARE": r = xf * xf
                elif op == "LOG1P": r = np.log1p(xf)
                elif op == "EXPM1": r = np.expm1(xf)
                elif op == "DIGAMMA":
                    from scipy.special import digamma
                    r = digamma(xf.astype(np.float64)).astype(np.float32)
                elif op == "ADD": r = xf + yf
                elif op == "SUB": r = xf - yf
                elif op == "MUL": r = xf * yf
                elif op == "DIV": r = xf / yf
                elif op == "MIN": r = np.minimum(xf, yf)
                elif op == "MAX": r = np.maximum(xf, yf)
                elif op == "POW": r = np.power(xf, yf)
                elif op == "HYPOT": r = np.hypot(xf, yf)
                elif op == "FMOD": r = np.fmod(xf, yf)
                elif op == "MOD": r = np.mod(xf, yf)
                elif op == "FDIV": r = np.floor(xf / yf)
                elif op == "SAFELOG": r = np.where(xf > 0, np.log(np.where(xf > 0, xf, one)), -np.inf) + 0 * yf
                elif op == "LT": r = (xf < yf).astype(np.int32)
                elif op == "LE": r = (xf <= yf).astype(np.int32)
                elif op == "GT": r = (xf > yf).astype(np.int32)
                elif op == "GE": r = (xf >= yf).astype(np.int32)
                elif op == "EQ": r = (xf == yf).astype(np.int32)
                elif op == "NE": r = (xf != yf).astype(np.int32)
                elif op == "MINW": r = np.where(xf < yf, one, np.where(xf == yf, np.float32(.5), np.float32(0)))
                elif op == "MAXW": r = np.where(xf > yf, one, np.where(xf == yf, np.float32(.5), np.float32(0)))
                elif op == "SELECT": r = np.where(x != 0, y, z)
                elif op == "FMA": r = (xf * yf) + zf
                elif op == "LERP": r = (xf * yf) + ((one - xf) * zf)
                elif op == "INEG": r = -xi
                elif op == "IABS": r = np.abs(xi)
                elif op == "ISGN": r = np.sign(xi)
                elif op == "IADD": r = xi + yi
                elif op == "ISUB": r = xi - yi
                elif op == "IMUL": r = xi * yi
                elif op == "IMIN": r = np.minimum(xi, yi)
                elif op == "IMAX": r = np.maximum(xi, yi)
                elif op == "IFDIV": r = np.where(yi == 0, 0, np.floor_divide(xi, np.where(yi == 0, 1, yi)))
                elif op == "IMOD": r = np.where(yi == 0, 0, np.mod(xi, np.where(yi == 0, 1, yi)))
                elif op == "IFMOD": r = np.where(yi == 0, 0, np.fmod(xi, np.where(yi == 0, 1, yi)))
                elif op == "IPOW": r = np.power(xi, np.maximum(yi, 0))
                elif op == "ILT": r = (xi < yi).astype(np.int32)
                elif op == "ILE": r = (xi <= yi).astype(np.int32)
                elif op == "IGT": r = (xi > yi).astype(np.int32)
                elif op == "IGE": r = (xi >= yi).astype(np.int32)
                elif op == "IEQ": r = (xi == yi).astype(np.int32)
                elif op == "INE": r = (xi != yi).astype(np.int32)
                elif op == "AND": r = xi & yi
                elif op == "OR": r = xi | yi
                elif op == "XOR": r = xi ^ yi
                elif op == "NOT": r = (xi == 0).astype(np.int32)
                elif op == "I2F": r = xi.astype(np.float32)
                elif op == "F2I": r = np.nan_to_num(np.trunc(xf), nan=0, posinf=2**31 - 1, neginf=-2**31).astype(np.int64).clip(-2**31, 2**31 - 1).astype(np.int32)
                elif op == "F2B": r = (xf != 0).astype(np.int32)
                elif op == "I2B": r = (xi != 0).astype(np.int32)
                else:
                    raise NotImplementedError(op)
            r = np.asarray(r)
            if r.dtype == np.float64:
                r = r.astype(np.float32)
            if r.dtype == np.int64:
                r = r.astype(np.int32)
            store(r)

"""Thread-per-rollout kernel generator: the lanes-per-rollout = 1 end of the mapping of core/codegen.py.

``core/codegen.py`` gives a rollout L lanes (lanes over object tuples) and keeps every value with at
most L elements in one register per lane.  With L = 1 a lane owns a whole rollout: a value of n
elements is n registers of that lane, every gather / broadcast / aggregation of the program is
resolved at generation time into plain register references (no shuffle, no table load, no idle
lane), and the only cross-lane operation left is the sum over the rollouts of a warp that the
plan gradient needs (QSUM, one xor-butterfly per gradient word).  The step of the MarsRover
adjoint drops from 332 SASS instructions per rollout-step to ~45 per rollout-step this way.

Trade-off: a warp now carries 32 rollouts, so the batch has to be large enough to fill the machine
(B >= ~32 * 2 * 592 rollouts for two warps per scheduler), and the live values of a step have to
fit the register file (what does not fit spills to lane-interleaved local memory).  The assembler's
cost model (core/program.py:_assemble) picks L = 1 only then.

Global memory layout is unchanged (fluent-major [W][B] blocks of n_f words per rollout): lane b reads
words b * n_f .. b * n_f + n_f - 1 of a block, so a warp still touches one contiguous run of
32 * n_f words per fluent.  Same program blob, same ``RkLaunch`` argument block, same C ABI
(rk_program_attach_cubin_ex carries the launch geometry: 128 threads per CTA, no shared memory).
"""
from __future__ import annotations

import os
from typing import Any, Dict, List

import numpy as np

from . import isa
from .codegen import _F1, _F2, _F2B, _FAST, _I1, _I2, _RED, KERNEL_NAME, CODEGEN_VERSION, _flit
from .program import Program

THREADS = 128


class _GenT:
    def __init__(self, prog: Program):
        assert prog.lanes == 1 and prog.rpw == 32
        self.p = prog
        self.ins = prog.meta["ins"]
        self.img = prog.meta["const_image"]
        self.mparam_off = prog.meta["mparam_off"]
        self.n_pro, self.n_step, self.n_epi = prog.n_ins
        self.fast = bool(prog.relaxed) and os.environ.get("RK_FAST_RELAXED", "1") != "0"
        self.fast_trig = self.fast or os.environ.get("RK_FAST_TRIG", "1") != "0"
        self.dtype: Dict[int, str] = {}
        self.size: Dict[int, int] = {}
        for rec in self.ins:
            for v in ([rec["dst"]] if rec["dst"] else []) + rec["srcs"]:
                if v["kind"] == "val":
                    self.dtype.setdefault(v["id"], v["dtype"])
                    self.size[v["id"]] = max(self.size.get(v["id"], 0), v["n"])
        self.lines: List[str] = []
        self.bases: Dict[int, str] = {}
        self.used_mparams = set()

    # ---- helpers --------------------------------------------------------------------------------
    @staticmethod
    def ctype(dt: str) -> str:
        return "float" if dt == "f" else "int"

    def emit(self, s: str):
        self.lines.append(s)

    def table(self, mp: int, n: int) -> np.ndarray:
        return self.img[mp: mp + n].astype(np.int64)

    def src_index(self, v, e: int, n_r: int) -> int:
        mp = v["map"]
        if mp == isa.MAP_IDENT:
            return e
        if mp == isa.MAP_BCAST:
            return 0
        return int(self.img[mp + e])

    def opnd(self, v, e: int, n_r: int, want: str) -> str:
        """C expression of element e of operand v, typed want ('f' float / 'i' int)."""
        want_c = "f" if want == "f" else "i"
        j = self.src_index(v, e, n_r)
        if v["kind"] == "const":
            c = v["const"].reshape(-1)
            val = c[j if c.size > 1 else 0]
            if v["dtype"] == "f":
                if want_c == "f":
                    return _flit(float(val))
                return f"((int)0x{np.float32(val).view(np.uint32):08x}u)"
            if want_c == "f":
                return f"__uint_as_float(0x{np.uint32(np.int32(val).view(np.uint32)):08x}u)"
            return f"({int(val)})"
        if v["kind"] == "mparam":
            slot = v["base"] - self.mparam_off
            self.used_mparams.add(slot)
            ex = f"mp{slot}"
            return ex if want_c == "f" else f"(int)__float_as_uint({ex})"
        name = f"v{v['id']}_{j}"
        src_c = "f" if self.dtype[v["id"]] == "f" else "i"
        if src_c != want_c:
            return f"__uint_as_float((uint32_t)({name}))" if want_c == "f" else f"(int)__float_as_uint({name})"
        return name

    def dst_name(self, v, e: int) -> str:
        return f"v{v['id']}_{e}"

    def gbase(self, buf: int, kind: str, tstep: bool) -> str:
        """Per list (and per step) base pointer of a global buffer: 'b' per-rollout blocks,
        'u' uniform rows, 'g' per-warp gradient partial rows."""
        name = f"gp{buf}"
        if buf not in self.bases:
            tt = "t" if tstep else "0ll"
            if kind == "b":
                ex = f"{tt} * K.W[{buf}] * (long long)K.B"
            elif kind == "u":
                ex = f"{tt} * K.W[{buf}]"
            else:
                ex = f"(gwarp * (long long)K.T + {tt}) * K.W[{buf}]"
            self.bases[buf] = f"  uint32_t* const {name} = (uint32_t*)K.buf[{buf}] + {ex};"
        return name

    # ---- one instruction ----------------------------------------------------------------------
    def gen_ins(self, rec):
        op, n = rec["op"], rec["n"]
        dst, srcs = rec["dst"], rec["srcs"]
        self.emit(f"  // {op} n={n}")
        if op == "NOP":
            return
        if op in ("LDB", "STB", "LDU", "STG", "STERR", "ERRCHK", "QSUM"):
            return self.gen_special(rec)
        if op in _RED or op in ("RARGMAX", "RARGMIN", "RPRODX"):
            return self.gen_reduce(rec)
        dt_out = dst["dtype"]
        out_c = "f" if dt_out == "f" else "i"
        recip = False
        if op == "DIV" and self.fast and srcs[1]["kind"] == "const":
            recip = True
        if self.fast and op in _FAST:
            body, tin = _FAST[op], ("ff" if op in _F2 else "f")
        elif self.fast_trig and op in ("SIN", "COS"):
            body, tin = _FAST[op], "f"
        elif self.fast_trig and op == "TAN":
            body, tin = "rk_ftan_ieee(x)", "f"
        elif op in _F1:
            body, tin = _F1[op], "f"
        elif op in _F2:
            body, tin = _F2[op], "ff"
        elif op in _F2B:
            body, tin = f"((x {_F2B[op]} y) ? 1 : 0)", "ff"
        elif op in _I1:
            body, tin = _I1[op], "i"
        elif op in _I2:
            body, tin = _I2[op], "ii"
        elif op == "MOV":
            body, tin = "x", out_c
        elif op == "SELECT":
            body, tin = "((x != 0) ? y : z)", "i" + out_c + out_c
        elif op == "FMA":
            body, tin = "__fadd_rn(__fmul_rn(x, y), z)", "fff"
        elif op == "LERP":
            body, tin = "__fadd_rn(__fmul_rn(x, y), __fmul_rn(__fsub_rn(1.f, x), z))", "fff"
        elif op == "I2F":
            body, tin = "(float)x", "i"
        elif op == "F2I":
            body, tin = "(int)x", "f"
        elif op == "F2B":
            body, tin = "((x != 0.f) ? 1 : 0)", "f"
        else:
            raise NotImplementedError(f"codegen_tpr: op {op}")
        names = "xyz"
        # an instruction may overwrite one of its own operands (persistent values): evaluate every
        # element into temporaries first when the destination is also read through a gather
        alias = any(s["kind"] == "val" and s["id"] == dst["id"] and s["map"] not in (isa.MAP_IDENT,)
                    for s in srcs)
        tmp = []
        for e in range(n):
            if recip:
                c = srcs[1]["const"].reshape(-1)
                cv = c[self.src_index(srcs[1], e, n) if c.size > 1 else 0]
                with np.errstate(all="ignore"):
                    r = _flit(float(np.float32(1.0) / np.float32(cv)))
                expr = f"({self.opnd(srcs[0], e, n, 'f')} * {r})"
                stmt = f"{self.ctype(out_c)} r_ = {expr};"
            else:
                decl = " ".join(f"const {self.ctype(tk)} {names[k]} = {self.opnd(srcs[k], e, n, tk)};"
                                for k, tk in enumerate(tin))
                stmt = f"{decl} const {self.ctype(out_c)} r_ = {body};"
            if alias:
                self.emit(f"  {self.ctype(out_c)} a{dst['id']}_{e}; {{ {stmt} a{dst['id']}_{e} = r_; }}")
                tmp.append(e)
            else:
                self.emit(f"  {{ {stmt} {self.dst_name(dst, e)} = r_; }}")
        for e in tmp:
            self.emit(f"  {self.dst_name(dst, e)} = a{dst['id']}_{e};")

    # ---- reductions (CSR tables resolved here) --------------------------------------------------
    def gen_reduce(self, rec):
        op, n = rec["op"], rec["n"]
        dst, src = rec["dst"], rec["srcs"][0]
        tab = rec["tab"]
        ptr = np.asarray(tab["ptr"]).astype(np.int64)
        idx = np.asarray(tab["idx"]).astype(np.int64)

        def sv(j: int, want: str) -> str:
            return self.opnd({**src, "map": isa.MAP_IDENT}, int(j), src["n"], want)

        if op == "RPRODX":
            seg = np.asarray(tab["seg"]).astype(np.int64)
            done = set()
            partition = len(idx) == n and np.array_equal(np.sort(idx), np.arange(n))
            if partition:
                for s in range(len(ptr) - 1):
                    mem = [int(m) for m in idx[ptr[s]:ptr[s + 1]]]
                    k = len(mem)
                    if k == 0:
                        continue
                    # prefix / suffix products within the segment
                    self.emit(f"  {{ float pre_ = 1.f;")
                    for j, m in enumerate(mem):
                        self.emit(f"    {self.dst_name(dst, m)} = pre_; pre_ = pre_ * {sv(m, 'f')};")
                    self.emit("    float suf_ = 1.f;")
                    for j in range(k - 1, -1, -1):
                        m = mem[j]
                        self.emit(f"    {self.dst_name(dst, m)} = {self.dst_name(dst, m)} * suf_; "
                                  f"suf_ = suf_ * {sv(m, 'f')};")
                    self.emit("  }")
                    done.update(mem)
            for e in range(n):
                if e in done:
                    continue
                s = int(seg[e])
                terms = [sv(m, "f") for m in idx[ptr[s]:ptr[s + 1]] if int(m) != e]
                self.emit(f"  {self.dst_name(dst, e)} = " + (" * ".join(terms) if terms else "1.f") + ";")
            return
        if op in ("RARGMAX", "RARGMIN"):
            sgn = "" if op == "RARGMAX" else "-"
            for s in range(n):
                mem = idx[ptr[s]:ptr[s + 1]]
                self.emit("  { float best_ = -CUDART_INF_F; int acc_ = 0;")
                for k, m in enumerate(mem):
                    self.emit(f"    {{ const float t_ = {sgn}{sv(m, 'f')}; "
                              f"if (t_ > best_) {{ best_ = t_; acc_ = {k}; }} }}")
                self.emit(f"    {self.dst_name(dst, s)} = acc_; }}")
            return
        ct, init, upd = _RED[op]
        tin = "f" if ct == "float" else "i"
        for s in range(n):
            mem = idx[ptr[s]:ptr[s + 1]]
            if len(mem) == 0:
                self.emit(f"  {self.dst_name(dst, s)} = {init};")
                continue
            self.emit(f"  {{ {ct} acc = {init};")
            for m in mem:
                self.emit(f"    {{ const {ct} t = {sv(m, tin)}; {upd}; }}")
            self.emit(f"    {self.dst_name(dst, s)} = acc; }}")

    # ---- memory / cross-rollout / error instructions --------------------------------------------
    def gen_special(self, rec):
        op, n = rec["op"], rec["n"]
        dst, srcs = rec["dst"], rec["srcs"]
        buf, off = rec["buf"], rec["off"]
        tstep = bool(rec["flags"] & isa.F_TSTEP)
        if op == "QSUM":
            src = srcs[0]
            for e in range(n):
                x = self.opnd(src, e, n, "f")
                self.emit(f"  {{ float acc = valid ? {x} : 0.f; "
                          + " ".join(f"acc += __shfl_xor_sync(0xffffffffu, acc, {o}u);"
                                     for o in (16, 8, 4, 2, 1))
                          + f" {self.dst_name(dst, e)} = acc; }}")
            return
        if op == "LDB":
            gp = self.gbase(buf, "b", tstep)
            dt = self.dtype[dst["id"]]
            # lane b owns words b*n .. b*n + n - 1 of the fluent's [B][n] block
            self.emit(f"  {{ const uint32_t* const src_ = {gp} + ({off}u * Bu + bq * {n}u);")
            vec = 4 if n % 4 == 0 else (2 if n % 2 == 0 else 1)
            if vec > 1:
                # the block starts at a multiple of B words: B % 4 == 0 (host-checked via K.vec_ok)
                # makes lane b's run 16-byte (8-byte) aligned -> one wide load per 4 (2) words
                vt = "uint4" if vec == 4 else "uint2"
                comps = "xyzw"[:vec]
                self.emit(f"    if (vec_ok) {{")
                for e0 in range(0, n, vec):
                    self.emit(f"      {{ const {vt} w_ = ok{buf} ? __ldg(reinterpret_cast<const {vt}*>(src_ + {e0})) "
                              f": make_{vt}({', '.join(['0u'] * vec)});")
                    for c in range(vec):
                        self.emit(f"        {self.dst_name(dst, e0 + c)} = "
                                  + (f"__uint_as_float(w_.{comps[c]});" if dt == "f" else f"(int)w_.{comps[c]};"))
                    self.emit("      }")
                self.emit("    } else {")
            for e in range(n):
                ld = f"(ok{buf} ? __ldg(src_ + {e}) : 0u)"
                self.emit(f"      {self.dst_name(dst, e)} = "
                          + (f"__uint_as_float({ld});" if dt == "f" else f"(int){ld};"))
            if vec > 1:
                self.emit("    }")
            self.emit("  }")
            return
        if op == "STB":
            src = srcs[0]
            gp = self.gbase(buf, "b", tstep)
            dtc = "f" if (src["dtype"] if src["kind"] != "val" else self.dtype[src["id"]]) == "f" else "i"
            self.emit(f"  if (ok{buf}) {{ uint32_t* const dst_ = {gp} + ({off}u * Bu + bq * {n}u);")
            vec = 4 if n % 4 == 0 else (2 if n % 2 == 0 else 1)
            words = []
            for e in range(n):
                val = self.opnd(src, e, n, dtc)
                words.append(f"__float_as_uint({val})" if dtc == "f" else f"(uint32_t)({val})")
            if vec > 1:
                vt = "uint4" if vec == 4 else "uint2"
                self.emit("    if (vec_ok) {")
                for e0 in range(0, n, vec):
                    self.emit(f"      *reinterpret_cast<{vt}*>(dst_ + {e0}) = make_{vt}("
                              + ", ".join(words[e0:e0 + vec]) + ");")
                self.emit("    } else {")
            for e in range(n):
                self.emit(f"      dst_[{e}] = {words[e]};")
            if vec > 1:
                self.emit("    }")
            self.emit("  }")
            return
        if op == "LDU":
            gp = self.gbase(buf, "u", tstep)
            dt = self.dtype[dst["id"]]
            for e in range(n):
                ld = f"(has{buf} ? __ldg({gp} + {off + e}u) : 0u)"
                self.emit(f"  {self.dst_name(dst, e)} = "
                          + (f"__uint_as_float({ld});" if dt == "f" else f"(int){ld};"))
            return
        if op == "STG":
            src = srcs[0]
            gp = self.gbase(buf, "g", tstep)
            # every lane holds the warp sums: lane e stores word e (coalesced), 32 words per pass
            for e0 in range(0, n, 32):
                m = min(32, n - e0)
                sel = self.opnd(src, e0, n, "f")
                for e in range(e0 + 1, e0 + m):
                    sel = f"(lane == {e - e0}u ? {self.opnd(src, e, n, 'f')} : {sel})"
                self.emit(f"  if (has{buf} && lane < {m}u) {gp}[{off + e0}u + lane] = __float_as_uint({sel});")
            return
        if op == "ERRCHK":
            src = srcs[0]
            bit = rec["aux0"]
            terms = " | ".join(f"({self.opnd(src, e, n, 'i')} != 0)" for e in range(n))
            self.emit(f"  if ({terms}) err_ |= {1 << bit}u;")
            return
        if op == "STERR":
            gp = self.gbase(buf, "b", tstep)
            self.emit(f"  if (ok{buf}) {gp}[bq] = err_ | {self.p.static_err}u;")
            self.emit("  err_ = 0u;")
            return
        raise NotImplementedError(op)

    # ---- whole kernel ---------------------------------------------------------------------------
    def _min_blocks(self) -> str:
        return _min_blocks_env(self.p.kind)

    def source(self) -> str:
        p = self.p
        lists = [self.ins[:self.n_pro], self.ins[self.n_pro:self.n_pro + self.n_step],
                 self.ins[self.n_pro + self.n_step:]]
        bodies = []
        for lst in lists:
            self.lines, self.bases = [], {}
            for rec in lst:
                self.gen_ins(rec)
            bodies.append("\n".join(list(self.bases.values()) + self.lines))
        decl = []
        for vid in sorted(self.size):
            ct = self.ctype(self.dtype[vid])
            zero = "0.f" if ct == "float" else "0"
            decl.append("  " + ct + " " + ", ".join(f"v{vid}_{e} = {zero}" for e in range(self.size[vid])) + ";")
        mps = []
        defaults = p.mparam_defaults
        for slot in sorted(self.used_mparams):
            mps.append(f"  const float mp{slot} = (K.mparams != nullptr) ? K.mparams[{slot}] : "
                       f"{_flit(float(defaults[slot]))};")
        hoist = [f"  const bool has{b} = K.buf[{b}] != nullptr; const bool ok{b} = has{b} && valid; "
                 f"(void)has{b}; (void)ok{b};" for b in range(len(isa.BUFFERS))]
        backward = p.kind == "backward"
        loop = "for (long long t = (long long)K.T - 1; t >= 0; --t)" if backward else \
            "for (long long t = 0; t < (long long)K.T; ++t)"
        nvals = sum(self.size.values())
        return f"""// GENERATED by pyrddlgym_jax_b200/core/codegen_tpr.py (version {CODEGEN_VERSION}) -- do not edit.
// thread-per-rollout kernel: kind={p.kind} relaxed={p.relaxed} S={p.S} A={p.A} N={p.N}
//          instructions={p.n_ins} scalar values={nvals}
#include "rk_launch.h"
#include "rk_device.cuh"

extern "C" __global__ void __launch_bounds__({THREADS}{self._min_blocks()}) {KERNEL_NAME}(const RkLaunch K) {{
  const uint32_t lane = threadIdx.x & 31u;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // this lane's rollout
  const long long gwarp = b >> 5;
  if ((gwarp << 5) >= (long long)K.B) return;                   // whole warp past the batch
  const bool valid = b < (long long)K.B;
  const uint32_t Bu = (uint32_t)K.B;                            // host checks W * B < 2^31
  const uint32_t bq = valid ? (uint32_t)b : 0u;
  const bool vec_ok = (Bu & 3u) == 0u;
  uint32_t err_ = 0u;
  (void)lane; (void)gwarp; (void)bq; (void)vec_ok; err_ |= 0u;
{chr(10).join(hoist)}
{chr(10).join(mps)}
{chr(10).join(decl)}
  {{ const long long t = 0; (void)t;
{bodies[0]}
  }}
  {loop} {{
{bodies[1]}
  }}
  {{ const long long t = 0; (void)t;
{bodies[2]}
  }}
}}
"""


def generate_source(prog: Program) -> str:
    return _GenT(prog).source()


def _min_blocks_env(kind: str) -> str:
    """A/B switch RK_TPR_MINBLOCKS (optionally 'forward=a,backward=b'): resident CTAs per SM the register
    allocation of a thread-per-rollout kernel has to leave room for (second __launch_bounds__ argument)."""
    raw = os.environ.get("RK_TPR_MINBLOCKS", "")
    if not raw:
        return ""
    val = raw
    if "=" in raw:
        val = dict(kv.split("=") for kv in raw.split(",")).get(kind, "")
    return f", {int(val)}" if val else ""

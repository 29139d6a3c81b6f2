#!/usr/bin/env python
"""Generate csrc/generated/spec_kinematics.cuh: straight-line FK / Jacobian code specialised
to one kinematic tree (the Panda chain of assets/panda_shelf_kinematic.xml).

Why: the generic kernels read the tree from __constant__ memory and multiply through every
3x3 link rotation (about 885 FLOP per DLS iteration).  The Panda's link twists are +-90 degree
rotations about x, most link offsets have one non-zero component, the base frame is axis
aligned and the EE site lies on joint 7's axis, so most of those products are by 0 or +-1.
SURVEY.md section 8(d) fixes the algorithmic cost at 500 FLOP/iteration *assuming* that
structure is exploited; this generator is how the build exploits it without hand-deriving
Panda-specific formulas (nothing here is Panda specific: any 7-hinge tree works).

How: a tiny expression compiler with *numeric fingerprints*.  Every scalar is a node that
carries its value at K random joint configurations (FP64).  After each operation the result's
fingerprint is checked: all ~0 -> the constant 0, all equal -> a literal constant, equal (or
opposite) to an existing node -> that node is reused with a sign.  This finds structural
zeros (e.g. the whole 7th Jacobian column), permutation-only rotations and common
subexpressions without symbolic algebra.  The emitted code is templated on the scalar type
and uses only * + - (nvcc contracts to FMA).

Run:  python tools/gen_spec_kinematics.py   (done by __graft_entry__.build(); output committed)
"""

from __future__ import annotations

import argparse
import hashlib
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mujoco_panda_pnp_b200.tree import KinematicTree  # noqa: E402

K_SAMPLES = 24
TOL = 1e-11


class Node:
    __slots__ = ("kind", "name", "const", "fp", "expr")

    def __init__(self, kind, name=None, const=None, fp=None, expr=None):
        self.kind, self.name, self.const, self.fp, self.expr = kind, name, const, fp, expr


class Signed:
    """A node reference with a sign: the unit every expression works on."""

    __slots__ = ("node", "sign")

    def __init__(self, node, sign=1):
        self.node, self.sign = node, sign

    @property
    def fp(self):
        return self.sign * self.node.fp

    @property
    def is_const(self):
        return self.node.kind == "const"

    @property
    def cval(self):
        return self.sign * self.node.const


class Compiler:
    def __init__(self, rng):
        self.rng = rng
        self.lines = []
        self.nodes = []  # runtime nodes for fingerprint CSE
        self.n_tmp = 0

    # --- leaves -----------------------------------------------------------------------
    def const(self, v):
        v = float(v)
        return Signed(Node("const", const=v, fp=np.full(K_SAMPLES, v)))

    def input(self, name, fp):
        n = Node("var", name=name, fp=np.asarray(fp, dtype=np.float64))
        self.nodes.append(n)
        return Signed(n)

    # --- helpers ----------------------------------------------------------------------
    def _canon(self, fp, expr_fn):
        """Fold the fingerprint to a constant / existing node, else emit a new temporary."""
        if np.all(np.abs(fp) < TOL):
            return self.const(0.0)
        if np.all(np.abs(fp - fp[0]) < TOL):
            v = fp[0]
            for snap in (1.0, -1.0, 0.5, -0.5, 2.0, -2.0):
                if abs(v - snap) < TOL:
                    v = snap
            return self.const(v)
        scale = max(1.0, float(np.max(np.abs(fp))))
        for n in self.nodes:
            if np.all(np.abs(n.fp - fp) < TOL * scale):
                return Signed(n, 1)
            if np.all(np.abs(n.fp + fp) < TOL * scale):
                return Signed(n, -1)
        name = f"t{self.n_tmp}"
        self.n_tmp += 1
        expr, op, deps, vexpr = expr_fn()
        self.lines.append(dict(name=name, expr=expr, op=op, deps=deps, vexpr=vexpr))
        n = Node("var", name=name, fp=fp)
        self.nodes.append(n)
        return Signed(n)

    @staticmethod
    def _lit(v):
        return f"T({v!r})"

    def ref(self, s, leading=True):
        """Source text of a signed operand."""
        if s.is_const:
            v = s.cval
            return self._lit(v) if leading or v >= 0 else f"({self._lit(v)})"
        return s.node.name if s.sign > 0 else (f"-{s.node.name}" if leading else f"(-{s.node.name})")

    def vref(self, s):
        """Operand text for the function-call ("_v") rendering: signs become pnp_neg()."""
        if s.is_const:
            return self._lit(s.cval)
        return s.node.name if s.sign > 0 else f"pnp_neg({s.node.name})"

    # --- arithmetic -------------------------------------------------------------------
    def mul(self, a, b):
        fp = a.fp * b.fp
        if a.is_const and b.is_const:
            return self.const(a.cval * b.cval)
        for x, y in ((a, b), (b, a)):
            if x.is_const:
                if x.cval == 0.0:
                    return self.const(0.0)
                if x.cval == 1.0:
                    return Signed(y.node, y.sign)
                if x.cval == -1.0:
                    return Signed(y.node, -y.sign)

        def emit():
            sign = a.sign * b.sign if not (a.is_const or b.is_const) else (b.sign if a.is_const else a.sign)
            ta = self._lit(a.cval) if a.is_const else a.node.name
            tb = self._lit(b.cval) if b.is_const else b.node.name
            # a constant carries its own sign in the literal
            if a.is_const or b.is_const:
                k, v = (a, b) if a.is_const else (b, a)
                vx = f"pnp_mul({v.node.name}, {self._lit(k.cval * v.sign)})"
            else:
                vx = f"pnp_mul({'pnp_neg(' + a.node.name + ')' if sign < 0 else a.node.name}, {b.node.name})"
            return f"{'-' if sign < 0 else ''}{ta} * {tb}", "mul", self._deps(a, b), vx

        return self._canon(fp, emit)

    def add(self, a, b):
        if a.is_const and b.is_const:
            return self.const(a.cval + b.cval)
        if a.is_const and a.cval == 0.0:
            return Signed(b.node, b.sign)
        if b.is_const and b.cval == 0.0:
            return Signed(a.node, a.sign)
        fp = a.fp + b.fp

        def emit():
            return (f"{self.ref(a)} + {self.ref(b, leading=False)}", "add", self._deps(a, b),
                    f"pnp_add({self.vref(a)}, {self.vref(b)})")

        return self._canon(fp, emit)

    def sub(self, a, b):
        return self.add(a, Signed(b.node, -b.sign) if not b.is_const else self.const(-b.cval))

    def fma(self, a, b, c):
        """a*b + c, emitted as one expression so nvcc contracts it."""
        if (a.is_const and a.cval == 0.0) or (b.is_const and b.cval == 0.0):
            return Signed(c.node, c.sign) if not c.is_const else self.const(c.cval)
        if c.is_const and c.cval == 0.0:
            return self.mul(a, b)
        if (a.is_const and abs(a.cval) == 1.0) or (b.is_const and abs(b.cval) == 1.0) or (a.is_const and b.is_const):
            return self.add(self.mul(a, b), c)
        fp = a.fp * b.fp + c.fp

        def emit():
            if a.is_const or b.is_const:
                k, v = (a, b) if a.is_const else (b, a)
                coef = k.cval * v.sign
                prod = f"{self._lit(coef)} * {v.node.name}"
                vx = f"pnp_fma({v.node.name}, {self._lit(coef)}, {self.vref(c)})"
            else:
                neg = a.sign * b.sign < 0
                prod = f"{'-' if neg else ''}{a.node.name} * {b.node.name}"
                vx = f"pnp_fma({'pnp_neg(' + a.node.name + ')' if neg else a.node.name}, {b.node.name}, {self.vref(c)})"
            return f"{prod} + {self.ref(c, leading=False)}", "fma", self._deps(a, b, c), vx

        return self._canon(fp, emit)

    def dot(self, xs, ys, init=None):
        acc = init if init is not None else self.const(0.0)
        # put constant-free products last so the first term can start the chain
        for x, y in zip(xs, ys):
            acc = self.fma(x, y, acc)
        return acc

    def store(self, target, s):
        self.lines.append(dict(name=None, expr=f"  {target} = {self.ref(s)};", op="store", deps=self._deps(s),
                               vexpr=f"  {target} = {self.vref(s)};"))

    @staticmethod
    def _deps(*ops):
        return [o.node.name for o in ops if not o.is_const]

    def finish(self):
        """Dead-code elimination from the stores backwards; returns (source lines, op stats)."""
        live, keep = set(), []
        for ln in reversed(self.lines):
            if ln["op"] == "store" or ln["name"] in live:
                keep.append(ln)
                live.update(ln["deps"])
        keep.reverse()
        stats = dict(mul=0, add=0, fma=0)
        out, vout = [], []
        for ln in keep:
            if ln["op"] == "store":
                out.append(ln["expr"])
                vout.append(ln["vexpr"])
            else:
                stats[ln["op"]] += 1
                out.append(f"  const T {ln['name']} = {ln['expr']};")
                vout.append(f"  const T {ln['name']} = {ln['vexpr']};")
        return out, stats, vout


def build_chain(cp: Compiler, tree: KinematicTree, qs: np.ndarray, want_rot: bool, joint1_frame: bool = False,
                fold_tail: bool = False):
    """Emit FK down the chain; returns (p_ee, R_ee or None, anchors, axes).

    joint1_frame: everything is expressed in the frame that joint 1 carries, A_0 Rz(q_1) - i.e. the chain starts at
    the identity with s_1 = 0, c_1 = 1 and the caller rotates the target into that frame (4 operations per pass).
    The DLS step J^T (J J^T + lam I)^-1 e does not depend on the frame J and e share, and in this one the first
    joint's rotation drops out of every product down the chain (Panda: 85 instead of 107 operations for FK + Jp,
    one more structural zero in Jp).

    fold_tail: the last joint L whose rotation moves the EE site is not multiplied into the frame; the (constant) offset
    from its origin to the site is rotated by q_L instead - w = Rz(q_L) v, p = p_L + R_L w: 4 + 9 operations where the
    frame product and the offsets cost 12 + 6.  Only valid when no later joint moves the site (its Jacobian columns are
    then structurally zero and are emitted as such); checked numerically against the unfolded chain."""
    s = [cp.input(f"s[{i}]", np.sin(qs[:, i] - tree.qref[i])) for i in range(7)]
    c = [cp.input(f"c[{i}]", np.cos(qs[:, i] - tree.qref[i])) for i in range(7)]
    if joint1_frame:
        s[0], c[0] = cp.const(0.0), cp.const(1.0)
    R = [[cp.const(1.0 if r == k else 0.0) for k in range(3)] for r in range(3)]
    p = [cp.const(0.0) for _ in range(3)]
    anchors, axes = [], []
    for i in range(7):
        pos = [cp.const(v) for v in (np.zeros(3) if joint1_frame and i == 0 else tree.link_pos[i])]
        rot = [[cp.const(v) for v in row] for row in (np.eye(3) if joint1_frame and i == 0 else tree.link_rot[i])]
        p = [cp.dot(R[r], pos, init=p[r]) for r in range(3)]
        R = [[cp.dot(R[r], [rot[k][j] for k in range(3)]) for j in range(3)] for r in range(3)]
        anchors.append(list(p))
        axes.append([R[r][2] for r in range(3)])
        need_xy = want_rot or i < 6 or np.any(np.abs(tree.ee_pos[:2]) > 0)
        last_xy = 5 if not np.any(np.abs(tree.ee_pos[:2]) > 0) else 6
        if fold_tail and not want_rot and i == last_xy:
            # offset from this joint's origin to the site in the frame after its rotation: later fixed transforms and
            # later joints' rotations folded from the end (constants if those joints cannot move the site)
            tail = [cp.const(v) for v in tree.ee_pos]
            for j in range(6, i, -1):
                tail = [cp.sub(cp.mul(c[j], tail[0]), cp.mul(s[j], tail[1])), cp.add(cp.mul(s[j], tail[0]), cp.mul(c[j], tail[1])), tail[2]]
                rot_j = [[cp.const(v) for v in row] for row in tree.link_rot[j]]
                tail = [cp.dot(rot_j[r], tail, init=cp.const(tree.link_pos[j][r])) for r in range(3)]
            if all(t.is_const for t in tail):
                sm = Signed(s[i].node, -s[i].sign)
                w = [cp.fma(c[i], tail[0], cp.mul(sm, tail[1])), cp.fma(s[i], tail[0], cp.mul(c[i], tail[1])), tail[2]]
                p_fold = [cp.dot(R[r], w, init=p[r]) for r in range(3)]
                for j in range(i + 1, 7):  # joints that cannot move the site: rel = 0 -> zero columns
                    anchors.append(list(p_fold))
                    axes.append([cp.const(0.0), cp.const(0.0), cp.const(1.0)])
                return p_fold, None, anchors, axes
        if need_xy:
            # emission order groups the three rows of each product so that consecutive instructions
            # share s_i (then c_i) in the same operand slot: on the packed f32x2 path the register file
            # delivers two 64-bit operands per instruction slot (a third costs an extra cycle unless it
            # sits in the operand-reuse cache, tools/microbench/fp32x2_operands.cu)
            mx = [cp.mul(s[i], R[r][1]) for r in range(3)]
            my = [cp.mul(cp.const(-s[i].cval) if s[i].is_const else Signed(s[i].node, -s[i].sign), R[r][0]) for r in range(3)]
            newx = [cp.fma(c[i], R[r][0], mx[r]) for r in range(3)]
            newy = [cp.fma(c[i], R[r][1], my[r]) for r in range(3)]
            R = [[newx[r], newy[r], R[r][2]] for r in range(3)]
    ee = [cp.const(v) for v in tree.ee_pos]
    p_ee = [cp.dot(R[r], ee, init=p[r]) for r in range(3)]
    R_ee = None
    if want_rot:
        er = [[cp.const(v) for v in row] for row in tree.ee_rot]
        R_ee = [[cp.dot(R[r], [er[k][j] for k in range(3)]) for j in range(3)] for r in range(3)]
    return p_ee, R_ee, anchors, axes


def cross(cp, a, b):
    return [
        cp.sub(cp.mul(a[1], b[2]), cp.mul(a[2], b[1])),
        cp.sub(cp.mul(a[2], b[0]), cp.mul(a[0], b[2])),
        cp.sub(cp.mul(a[0], b[1]), cp.mul(a[1], b[0])),
    ]


def cross_fused(cp, a, b):
    """a x b with each component as mul + fma (one rounding less, matches nvcc contraction)."""
    def comp(i, j):
        return cp.fma(a[i], b[j], cp.mul(Signed(a[j].node, -a[j].sign) if not a[j].is_const else cp.const(-a[j].cval), b[i]))
    return [comp(1, 2), comp(2, 0), comp(0, 1)]


def gen_function(tree, rng, name, want_jacp, want_full, joint1_frame=False, fold_tail=False):
    qs = rng.uniform(-3.0, 3.0, size=(K_SAMPLES, 7))
    cp = Compiler(rng)
    p_ee, R_ee, anchors, axes = build_chain(cp, tree, qs, want_rot=want_full, joint1_frame=joint1_frame, fold_tail=fold_tail)
    if fold_tail:  # the folded chain must be the same function of q as the plain one
        ref = build_chain(Compiler(rng), tree, qs, want_rot=want_full, joint1_frame=joint1_frame)
        assert all(np.allclose(a.fp, b.fp, atol=1e-9) for a, b in zip(p_ee, ref[0])), "fold_tail changed the site position"
    out_lines = []
    jp_zero = np.ones((3, 7), dtype=bool)
    jr_zero = np.ones((3, 7), dtype=bool)
    for r in range(3):
        cp.store(f"p[{r}]", p_ee[r])
    if want_jacp or want_full:
        for j in range(7):
            rel = [cp.sub(p_ee[r], anchors[j][r]) for r in range(3)]
            col = cross_fused(cp, axes[j], rel)
            for r in range(3):
                if not (col[r].is_const and col[r].cval == 0.0):
                    jp_zero[r, j] = False
                    cp.store(f"J[{r * 7 + j}]", col[r])
    if want_full:
        for j in range(7):
            for r in range(3):
                if not (axes[j][r].is_const and axes[j][r].cval == 0.0):
                    jr_zero[r, j] = False
                cp.store(f"J[{(3 + r) * 7 + j}]", axes[j][r])
        for r in range(3):
            for k in range(3):
                cp.store(f"R[{r * 3 + k}]", R_ee[r][k])
    sig = "const T* __restrict__ s, const T* __restrict__ c, T* __restrict__ p"
    if want_jacp or want_full:
        sig += ", T* __restrict__ J"
    if want_full:
        sig += ", T* __restrict__ R"
    out_lines.append(f"template <typename T>\n__device__ __forceinline__ void {name}({sig}) {{")
    body, stats, vbody = cp.finish()
    out_lines.extend(body)
    out_lines.append("}")
    # the same straight-line program in function-call form (pnp_fma / pnp_mul / pnp_add / pnp_neg,
    # overloaded for float, double and the packed f32x2 pair type): explicit FMAs, no reliance on
    # contraction, so it also instantiates for types the compiler cannot contract (FFMA2).
    out_lines.append(f"template <typename T>\n__device__ __forceinline__ void {name}_v({sig}) {{")
    out_lines.extend(vbody)
    out_lines.append("}")
    return "\n".join(out_lines), stats, jp_zero, jr_zero


def tree_fingerprint(tree: KinematicTree) -> str:
    blob = b"".join(
        struct.pack(f"<{a.size}d", *np.asarray(a, dtype=np.float64).reshape(-1))
        for a in (tree.link_pos, tree.link_rot, tree.ee_pos, tree.ee_rot, tree.lower, tree.upper, tree.qref)
    )
    return hashlib.sha256(blob).hexdigest()


def carr(name, a, fmt="%r"):
    flat = np.asarray(a, dtype=np.float64).reshape(-1)
    return f"static constexpr double {name}[{flat.size}] = {{" + ", ".join(fmt % float(v) for v in flat) + "};"


def cfun(name, a):
    """constexpr selector usable from device code with a (compile-time unrolled) index."""
    flat = [float(v) for v in np.asarray(a, dtype=np.float64).reshape(-1)]
    chain = " : ".join(f"i == {k} ? T({v!r})" for k, v in enumerate(flat[:-1]))
    return (f"template <typename T>\n__host__ __device__ __forceinline__ constexpr T {name}(int i) {{\n"
            f"  return {chain} : T({flat[-1]!r});\n}}")


def generate(tree: KinematicTree, src_desc: str) -> str:
    snapped = tree.snapped()
    rng = np.random.default_rng(12345)
    f_pos, st_pos, _, _ = gen_function(snapped, rng, "spec_fk_pos", False, False)
    f_jac, st_jac, jp_zero, _ = gen_function(snapped, rng, "spec_fk_jacp", True, False)
    f_full, st_full, _, _ = gen_function(snapped, rng, "spec_fk_full", False, True)
    _, _, jp_zero_plain, _ = gen_function(snapped, np.random.default_rng(777), "unused", True, False, joint1_frame=True)
    f_j1, st_j1, jp_zero_j1, _ = gen_function(snapped, rng, "spec_fk_jacp_j1", True, False, joint1_frame=True, fold_tail=True)
    assert (jp_zero_j1 == jp_zero_plain).all(), "fold_tail zeroed a Jacobian column that the plain chain does not"

    # world <-> frame of joint 1's parent (A_0 = the first fixed transform of the canonical chain)
    def rigid(name, fn_doc, apply):
        cp = Compiler(rng)
        fp = rng.uniform(-2.0, 2.0, size=(3, K_SAMPLES))
        x = [cp.input(f"x[{r}]", fp[r]) for r in range(3)]
        for r, v in enumerate(apply(cp, x)):
            cp.store(f"y[{r}]", v)
        _, _, vbody = cp.finish()
        return (f"// {fn_doc}\ntemplate <typename T>\n__device__ __forceinline__ void {name}(const T* __restrict__ x, T* __restrict__ y) {{\n"
                + "\n".join(vbody) + "\n}")

    a0_pos, a0_rot = snapped.link_pos[0], snapped.link_rot[0]

    def to_base(cp, x):  # A_0.rot^T (x - A_0.pos)
        d = [cp.sub(x[r], cp.const(a0_pos[r])) for r in range(3)]
        return [cp.dot([cp.const(a0_rot[r][k]) for r in range(3)], d) for k in range(3)]

    def to_world(cp, x):  # A_0.pos + A_0.rot x
        return [cp.dot([cp.const(a0_rot[r][k]) for k in range(3)], x, init=cp.const(a0_pos[r])) for r in range(3)]

    def to_base_vec(cp, x):  # A_0.rot^T x (a difference of two points)
        return [cp.dot([cp.const(a0_rot[r][k]) for r in range(3)], x) for k in range(3)]

    f_rigid = "\n\n".join([
        rigid("spec_world_to_base_v", "a world point in the frame of joint 1's parent: A_0.rot^T (x - A_0.pos)", to_base),
        rigid("spec_base_to_world_v", "and back: A_0.pos + A_0.rot x", to_world),
        rigid("spec_world_to_base_vec_v", "a world displacement in that frame: A_0.rot^T x", to_base_vec),
    ])

    # A = Jp Jp^T (upper triangle, 6 unique) and dq = Jp^T y skipping structural zeros
    jjt = ["template <typename T>\n__device__ __forceinline__ void spec_jjt(const T* __restrict__ J, T* __restrict__ A) {"]
    pairs = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    flops_jjt = 0
    for k, (r, s_) in enumerate(pairs):
        terms = [f"J[{r * 7 + j}] * J[{s_ * 7 + j}]" for j in range(7) if not (jp_zero[r, j] or jp_zero[s_, j])]
        flops_jjt += 2 * len(terms) - 1 if terms else 0
        jjt.append(f"  A[{k}] = {' + '.join(terms) if terms else 'T(0)'};")
    jjt.append("}")
    jjt.append("template <typename T>\n__device__ __forceinline__ void spec_jjt_v(const T* __restrict__ J, T* __restrict__ A) {")
    # column-major accumulation: the products of one Jacobian column are emitted back to back and share
    # J[r][j] in the first operand slot (operand-reuse cache, see build_chain)
    started = [False] * 6
    for j in range(7):
        for k, (r, s_) in enumerate(pairs):
            if jp_zero[r, j] or jp_zero[s_, j]:
                continue
            a_, b_ = f"J[{r * 7 + j}]", f"J[{s_ * 7 + j}]"
            jjt.append(f"  {'T ' if not started[k] else ''}a{k} = " + (f"pnp_fma({a_}, {b_}, a{k});" if started[k] else f"pnp_mul({a_}, {b_});"))
            started[k] = True
    for k in range(6):
        jjt.append(f"  A[{k}] = {'a%d' % k if started[k] else 'T(0.0)'};")
    jjt.append("}")
    # A = Jp Jp^T + lam I with lam as the start value of the three diagonal sums (one FMA instead of a multiply and,
    # later, an add per diagonal entry: three packed instructions fewer per DLS pass)
    jjt.append("template <typename T>\n__device__ __forceinline__ void spec_jjt_damped_v(const T* __restrict__ J, const T lam, T* __restrict__ A) {")
    started = [False] * 6
    for j in range(7):
        for k, (r, s_) in enumerate(pairs):
            if jp_zero[r, j] or jp_zero[s_, j]:
                continue
            a_, b_ = f"J[{r * 7 + j}]", f"J[{s_ * 7 + j}]"
            if started[k]:
                jjt.append(f"  a{k} = pnp_fma({a_}, {b_}, a{k});")
            elif r == s_:
                jjt.append(f"  T a{k} = pnp_fma({a_}, {b_}, lam);")
            else:
                jjt.append(f"  T a{k} = pnp_mul({a_}, {b_});")
            started[k] = True
    for k, (r, s_) in enumerate(pairs):
        jjt.append(f"  A[{k}] = {'a%d' % k if started[k] else ('lam' if r == s_ else 'T(0.0)')};")
    jjt.append("}")
    jty = ["template <typename T>\n__device__ __forceinline__ void spec_jty(const T* __restrict__ J, const T* __restrict__ y, T* __restrict__ dq) {"]
    flops_jty = 0
    for j in range(7):
        terms = [f"J[{r * 7 + j}] * y[{r}]" for r in range(3) if not jp_zero[r, j]]
        flops_jty += 2 * len(terms) - 1 if terms else 0
        jty.append(f"  dq[{j}] = {' + '.join(terms) if terms else 'T(0)'};")
    jty.append("}")
    jty.append("template <typename T>\n__device__ __forceinline__ void spec_jty_v(const T* __restrict__ J, const T* __restrict__ y, T* __restrict__ dq) {")
    # row-major accumulation: y[r] stays in the second operand slot across the seven joints
    started = [False] * 7
    for r in range(3):
        for j in range(7):
            if jp_zero[r, j]:
                continue
            jty.append(f"  {'T ' if not started[j] else ''}d{j} = " + (f"pnp_fma(J[{r * 7 + j}], y[{r}], d{j});" if started[j] else f"pnp_mul(J[{r * 7 + j}], y[{r}]);"))
            started[j] = True
    for j in range(7):
        jty.append(f"  dq[{j}] = {'d%d' % j if started[j] else 'T(0.0)'};")
    jty.append("}")

    # the same two routines for the Jacobian of spec_fk_jacp_j1 (its own structural zeros)
    jjt.append("template <typename T>\n__device__ __forceinline__ void spec_jjt_damped_j1_v(const T* __restrict__ J, const T lam, T* __restrict__ A) {")
    started = [False] * 6
    for j in range(7):
        for k, (r, s_) in enumerate(pairs):
            if jp_zero_j1[r, j] or jp_zero_j1[s_, j]:
                continue
            a_, b_ = f"J[{r * 7 + j}]", f"J[{s_ * 7 + j}]"
            if started[k]:
                jjt.append(f"  a{k} = pnp_fma({a_}, {b_}, a{k});")
            elif r == s_:
                jjt.append(f"  T a{k} = pnp_fma({a_}, {b_}, lam);")
            else:
                jjt.append(f"  T a{k} = pnp_mul({a_}, {b_});")
            started[k] = True
    for k, (r, s_) in enumerate(pairs):
        jjt.append(f"  A[{k}] = {'a%d' % k if started[k] else ('lam' if r == s_ else 'T(0.0)')};")
    jjt.append("}")
    jty.append("template <typename T>\n__device__ __forceinline__ void spec_jty_j1_v(const T* __restrict__ J, const T* __restrict__ y, T* __restrict__ dq) {")
    started = [False] * 7
    for r in range(3):
        for j in range(7):
            if jp_zero_j1[r, j]:
                continue
            jty.append(f"  {'T ' if not started[j] else ''}d{j} = " + (f"pnp_fma(J[{r * 7 + j}], y[{r}], d{j});" if started[j] else f"pnp_mul(J[{r * 7 + j}], y[{r}]);"))
            started[j] = True
    for j in range(7):
        jty.append(f"  dq[{j}] = {'d%d' % j if started[j] else 'T(0.0)'};")
    jty.append("}")

    def flop(st):
        return st["mul"] + st["add"] + 2 * st["fma"]

    col_zero = " || ".join(f"j == {j}" for j in range(7) if jp_zero[:, j].all()) or "false"
    mask_rows = ", ".join("{" + ", ".join("true" if z else "false" for z in row) + "}" for row in jp_zero)
    hdr = f"""// GENERATED by tools/gen_spec_kinematics.py - do not edit.
// Source tree: {src_desc}
// Straight-line kinematics specialised to that tree (structural zeros / +-1 folded by numeric
// fingerprinting, see the generator's docstring).  Operation counts (mul + add + 2*fma):
//   spec_fk_pos   {flop(st_pos):4d} FLOP   ({st_pos})
//   spec_fk_jacp  {flop(st_jac):4d} FLOP   ({st_jac})
//   spec_fk_full  {flop(st_full):4d} FLOP   ({st_full})
//   spec_fk_jacp_j1 {flop(st_j1):4d} FLOP ({st_j1}; FK + Jp in the frame joint 1 carries, what the FP32 IK kernels evaluate)
//   spec_jjt      {flops_jjt:4d} FLOP      spec_jty {flops_jty:4d} FLOP
#pragma once

namespace pnp_spec {{

static constexpr char kTreeSha256[] = "{tree_fingerprint(tree)}";
{carr('kLinkPos', tree.link_pos)}
{carr('kLinkRot', tree.link_rot)}
{carr('kEePos', tree.ee_pos)}
{carr('kEeRot', tree.ee_rot)}
{carr('kLower', tree.lower)}
{carr('kUpper', tree.upper)}
{carr('kQref', tree.qref)}
{cfun('spec_lower', tree.lower)}
{cfun('spec_upper', tree.upper)}
{cfun('spec_qref', tree.qref)}
// structurally-zero entries of the 3x7 position Jacobian
static constexpr bool kJpZero[3][7] = {{{mask_rows}}};
// joints whose whole position-Jacobian column is structurally zero (they cannot move the EE site)
__host__ __device__ __forceinline__ constexpr bool spec_jp_col_zero(int j) {{ return {col_zero}; }}

{f_pos}

{f_jac}

{f_full}

{f_j1}

{f_rigid}

{chr(10).join(jjt)}

{chr(10).join(jty)}

}}  // namespace pnp_spec
"""
    return hdr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mjcf", default=None, help="MJCF file (default: packaged kinematic asset)")
    ap.add_argument("--out", default=os.path.join(ROOT, "mujoco_panda_pnp_b200", "csrc", "generated", "spec_kinematics.cuh"))
    ap.add_argument("--check", action="store_true", help="exit 1 if the file on disk differs")
    args = ap.parse_args()
    tree = KinematicTree.from_mjcf(args.mjcf)
    desc = os.path.relpath(args.mjcf, ROOT) if args.mjcf else "mujoco_panda_pnp_b200/assets/panda_shelf_kinematic.xml"
    text = generate(tree, desc)
    if args.check:
        with open(args.out) as fh:
            sys.exit(0 if fh.read() == text else 1)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        fh.write(text)
    print(f"wrote {os.path.relpath(args.out, ROOT)} ({len(text.splitlines())} lines)")


if __name__ == "__main__":
    main()

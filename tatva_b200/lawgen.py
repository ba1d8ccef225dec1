"""Energy density -> CUDA constitutive law, by source-to-source automatic differentiation.

In the reference the density psi is USER code and JAX derives the residual (`jax.grad`) and the Hessian-vector product
(`jax.jvp(jax.grad(E))`, README.md:93, tatva/sparse/base.py:264) from it.  Here the user writes psi ONCE, on symbols:

    def psi(G, mu, lmbda):            # G: sympy Matrix (dofs_per_node x dim) = d u_c / d x_j at a quadrature point
        F = sympy.eye(3) + G
        J = F.det()
        return mu / 2 * ((F.T * F).trace() - 3 - 2 * sympy.log(J)) + lmbda / 2 * sympy.log(J) ** 2

and this module turns it into the `Mat` concept of csrc/common.cuh (`psi`, `first` = d psi / d G, `second` = its directional
derivative along dG) the way an AD system would — NOT by symbolic differentiation of the expanded expression (which swells:
2280 operations for the Mooney-Rivlin tangent) but on the straight-line program of psi:

    psi  --cse-->  straight-line program  --reverse sweep-->  P = d psi / d G  (~3-4 x the cost of psi)
                                          --forward (tangent) sweep over both-->  dP = (d P / d G) : dG  (~2-3 x that)

i.e. forward-over-reverse, exactly the composition `jax.jvp(jax.grad(.))` evaluates.  The emitted CUDA struct is compiled
at run time by NVRTC into `k_fused<El, UserLaw, MODE>` (csrc/user_law.cpp) and runs through the same fused kernels as the
built-in laws.  `emit_c` gives the same program as plain C for host-side checks.
"""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class _Prog:
    """SSA straight-line program: ops[i] = (dst, opcode, args); names are strings; literals are ('const', value)."""

    ops: list = field(default_factory=list)
    n: int = 0
    memo: dict = field(default_factory=dict)

    def new(self, op, *args):
        key = (op,) + args
        if key in self.memo:
            return self.memo[key]
        # algebraic simplifications with literal operands keep the tangent sweep from emitting dead arithmetic
        simp = _simplify(op, args)
        if simp is not None:
            return simp
        name = f"t{self.n}"
        self.n += 1
        self.ops.append((name, op, args))
        self.memo[key] = name
        return name


def _is_const(a):
    return isinstance(a, tuple) and a[0] == "const"


def _c(v):
    return ("const", float(v))


ZERO, ONE = _c(0.0), _c(1.0)


def _simplify(op, args):
    if op in ("add", "sub", "mul", "div") and all(_is_const(a) for a in args):
        a, b = args[0][1], args[1][1]
        return _c({"add": a + b, "sub": a - b, "mul": a * b, "div": a / b if b != 0 else float("nan")}[op])
    if op == "add":
        if args[0] == ZERO:
            return args[1]
        if args[1] == ZERO:
            return args[0]
    if op == "sub" and args[1] == ZERO:
        return args[0]
    if op == "mul":
        if ZERO in args:
            return ZERO
        if args[0] == ONE:
            return args[1]
        if args[1] == ONE:
            return args[0]
    if op == "div":
        if args[0] == ZERO:
            return ZERO
        if args[1] == ONE:
            return args[0]
    if op == "neg":
        if _is_const(args[0]):
            return _c(-args[0][1])
    return None


def _lower(expr, prog, env):
    """sympy expression -> SSA name (or literal), elementary operations only."""
    import sympy as sp

    if expr in env:
        return env[expr]
    if expr.is_Number:
        return _c(float(expr))
    if expr.is_Symbol:
        raise KeyError(f"unknown symbol {expr} in the energy density")
    if isinstance(expr, sp.Add):
        pos, neg = [], []
        for t in expr.args:
            c, rest = t.as_coeff_Mul()
            (neg if c.is_negative else pos).append((abs(c), rest))
        acc = None
        for c, rest in pos:
            v = _lower(c * rest, prog, env)
            acc = v if acc is None else prog.new("add", acc, v)
        for c, rest in neg:
            v = _lower(c * rest, prog, env)
            acc = prog.new("neg", v) if acc is None else prog.new("sub", acc, v)
        r = acc
    elif isinstance(expr, sp.Mul):
        num, den = [], []
        for t in expr.args:
            if isinstance(t, sp.Pow) and t.exp.is_Number and t.exp.is_negative:
                den.append(sp.Pow(t.base, -t.exp))
            else:
                num.append(t)
        acc = None
        for t in num:
            v = _lower(t, prog, env)
            acc = v if acc is None else prog.new("mul", acc, v)
        if acc is None:
            acc = ONE
        if den:
            d = None
            for t in den:
                v = _lower(t, prog, env)
                d = v if d is None else prog.new("mul", d, v)
            acc = prog.new("div", acc, d)
        r = acc
    elif isinstance(expr, sp.Pow):
        b = _lower(expr.base, prog, env)
        e = expr.exp
        if e.is_Integer and 2 <= int(e) <= 4:
            r = b
            for _ in range(int(e) - 1):
                r = prog.new("mul", r, b)
        elif e == sp.Rational(1, 2):
            r = prog.new("sqrt", b)
        elif e.is_Number:
            if e.is_negative:
                r = prog.new("div", ONE, _lower(sp.Pow(expr.base, -e), prog, env))
            elif e.is_Integer:
                r = prog.new("powc", b, _c(float(e)))
            else:
                # b ** k for a non-integer constant k (b > 0, e.g. J ** (-2/3)): exp(k log b).  Several powers of one base
                # share the logarithm, and exp / log cost a fraction of the general pow() with its special cases.
                r = prog.new("exp", prog.new("mul", _c(float(e)), prog.new("log", b)))
        else:
            r = prog.new("exp", prog.new("mul", _lower(e, prog, env), prog.new("log", b)))
    elif isinstance(expr, sp.log):
        r = prog.new("log", _lower(expr.args[0], prog, env))
    elif isinstance(expr, sp.exp):
        r = prog.new("exp", _lower(expr.args[0], prog, env))
    else:
        raise NotImplementedError(f"energy densities may use + - * / ** log exp sqrt; got {type(expr).__name__}")
    env[expr] = r
    return r


def _reverse(prog, out, inputs):
    """Append the reverse sweep to `prog`; returns {input name: adjoint name}."""
    adj = {out: ONE}

    def acc(v, t):
        if _is_const(v) or t == ZERO:
            return
        adj[v] = t if v not in adj else prog.new("add", adj[v], t)

    for name, op, a in list(reversed(prog.ops)):
        if name not in adj:
            continue
        b = adj[name]
        if op == "add":
            acc(a[0], b)
            acc(a[1], b)
        elif op == "sub":
            acc(a[0], b)
            acc(a[1], prog.new("neg", b))
        elif op == "neg":
            acc(a[0], prog.new("neg", b))
        elif op == "mul":
            acc(a[0], prog.new("mul", b, a[1]))
            acc(a[1], prog.new("mul", b, a[0]))
        elif op == "div":  # name = a0 / a1
            q = prog.new("div", b, a[1])
            acc(a[0], q)
            acc(a[1], prog.new("neg", prog.new("mul", q, name)))
        elif op == "log":
            acc(a[0], prog.new("div", b, a[0]))
        elif op == "exp":
            acc(a[0], prog.new("mul", b, name))
        elif op == "sqrt":
            acc(a[0], prog.new("div", b, prog.new("mul", _c(2.0), name)))
        elif op == "powc":  # name = a0 ** k  ->  k * name / a0
            acc(a[0], prog.new("mul", b, prog.new("div", prog.new("mul", a[1], name), a[0])))
        else:
            raise AssertionError(op)
    return {v: adj.get(v, ZERO) for v in inputs}


def _tangent(prog, seeds):
    """Forward (tangent) sweep over every op of `prog`; seeds: {input name: tangent name}.  Returns {name: tangent}."""
    tan = dict(seeds)

    def d(v):
        return ZERO if _is_const(v) else tan.get(v, ZERO)

    for name, op, a in list(prog.ops):
        if op == "add":
            t = prog.new("add", d(a[0]), d(a[1]))
        elif op == "sub":
            t = prog.new("sub", d(a[0]), d(a[1])) if d(a[0]) != ZERO else prog.new("neg", d(a[1]))
        elif op == "neg":
            t = prog.new("neg", d(a[0]))
        elif op == "mul":
            t = prog.new("add", prog.new("mul", d(a[0]), a[1]), prog.new("mul", a[0], d(a[1])))
        elif op == "div":  # (da - name * db) / b
            t = prog.new("div", prog.new("sub", d(a[0]), prog.new("mul", name, d(a[1]))), a[1])
        elif op == "log":
            t = prog.new("div", d(a[0]), a[0])
        elif op == "exp":
            t = prog.new("mul", d(a[0]), name)
        elif op == "sqrt":
            t = prog.new("div", d(a[0]), prog.new("mul", _c(2.0), name))
        elif op == "powc":
            t = prog.new("mul", d(a[0]), prog.new("div", prog.new("mul", a[1], name), a[0]))
        else:
            raise AssertionError(op)
        tan[name] = t
    return tan


def _live(ops, outs):
    need = {o for o in outs if not _is_const(o)}
    keep = []
    for name, op, a in reversed(ops):
        if name in need:
            keep.append((name, op, a))
            need.update(x for x in a if not _is_const(x))
    return list(reversed(keep))


def _fmt(a):
    return repr(a[1]) if _is_const(a) else a


_C_OPS = {"add": "{0} + {1}", "sub": "{0} - {1}", "mul": "{0} * {1}", "div": "{0} / {1}", "neg": "-{0}", "log": "log({0})", "exp": "exp({0})", "sqrt": "sqrt({0})", "powc": "pow({0}, {1})"}


def _emit(ops, indent="    "):
    return "".join(f"{indent}const double {n} = {_C_OPS[op].format(*[_fmt(x) for x in a])};\n" for n, op, a in ops)


@dataclass
class GeneratedLaw:
    """The three programs of a density and their operation counts."""

    dim: int
    dpn: int
    n_params: int
    uses_values: bool
    psi_ops: list
    first_ops: list
    second_ops: list
    psi_out: object
    first_out: dict  # input name -> SSA name of d psi / d input
    second_out: dict

    def op_counts(self):
        return dict(psi=len(self.psi_ops), first=len(self.first_ops), second=len(self.second_ops))

    # ---- code emission ----------------------------------------------------------------------------------
    def _inputs_decl(self, src, dsrc=None):
        out = ""
        for c in range(self.dpn):
            for j in range(self.dim):
                out += f"    const double G{c}{j} = {src}.G[{c}][{j}];\n"
                if dsrc:
                    out += f"    const double dG{c}{j} = {dsrc}.G[{c}][{j}];\n"
            if self.uses_values:
                out += f"    const double V{c} = {src}.val[{c}];\n"
                if dsrc:
                    out += f"    const double dV{c} = {dsrc}.val[{c}];\n"
        for k in range(self.n_params):
            out += f"    const double p{k} = prm[{k}];\n"
        return out

    def _store(self, outs, dst):
        s = ""
        for c in range(self.dpn):
            for j in range(self.dim):
                s += f"    {dst}.G[{c}][{j}] = {_fmt(outs[f'G{c}{j}'])};\n"
            if self.uses_values:
                s += f"    {dst}.val[{c}] = {_fmt(outs[f'V{c}'])};\n"
        return s

    def cuda_source(self, name="UserLaw"):
        """A struct with the `Mat` interface of csrc/common.cuh (what k_fused<El, Mat, MODE> expects)."""
        np_ = max(self.n_params, 1)
        return (
            f"struct {name} {{\n"
            f"  static constexpr int dim = {self.dim}, dpn = {self.dpn}, val_lo = {0 if self.uses_values else self.dpn}, n_params = {self.n_params};\n"
            f"  static constexpr bool needs_u_for_hvp = true;\n"
            f"  double prm[{np_}];\n"
            f"  struct Cache {{}};\n"
            f"  using S = tatva::QState<dpn, dim>;\n"
            f"  TATVA_HD void prepare(const S&, Cache&) const {{}}\n"
            f"  TATVA_HD double psi(const S& s, const Cache&) const {{\n{self._inputs_decl('s')}{_emit(self.psi_ops)}    return {_fmt(self.psi_out)};\n  }}\n"
            f"  TATVA_HD void first(const S& s, const Cache&, S& f) const {{\n{self._inputs_decl('s')}{_emit(self.first_ops)}{self._store(self.first_out, 'f')}  }}\n"
            f"  TATVA_HD void second(const S& s, const Cache&, const S& ds, S& f) const {{\n{self._inputs_decl('s', 'ds')}{_emit(self.second_ops)}{self._store(self.second_out, 'f')}  }}\n"
            f"}};\n"
        )

    def c_source(self):
        """Plain C twins for host-side checks: law_psi / law_first / law_second on flat arrays
        (G: dpn*dim row-major [, then dpn values]; prm: n_params)."""
        nin = self.dpn * self.dim + (self.dpn if self.uses_values else 0)

        def unpack(src, pre=""):
            s, k = "", 0
            for c in range(self.dpn):
                for j in range(self.dim):
                    s += f"    const double {pre}G{c}{j} = {src}[{k}];\n"
                    k += 1
            if self.uses_values:
                for c in range(self.dpn):
                    s += f"    const double {pre}V{c} = {src}[{k}];\n"
                    k += 1
            return s

        def pack(outs):
            names = [f"G{c}{j}" for c in range(self.dpn) for j in range(self.dim)] + ([f"V{c}" for c in range(self.dpn)] if self.uses_values else [])
            return "".join(f"    out[{k}] = {_fmt(outs[n])};\n" for k, n in enumerate(names))

        prm = "".join(f"    const double p{k} = prm[{k}];\n" for k in range(self.n_params))
        return (
            "#include <math.h>\n"
            f"int law_n_inputs(void) {{ return {nin}; }}\n"
            f"double law_psi(const double* in, const double* prm) {{\n{unpack('in')}{prm}{_emit(self.psi_ops)}    return {_fmt(self.psi_out)};\n}}\n"
            f"void law_first(const double* in, const double* prm, double* out) {{\n{unpack('in')}{prm}{_emit(self.first_ops)}{pack(self.first_out)}}}\n"
            f"void law_second(const double* in, const double* din, const double* prm, double* out) {{\n{unpack('in')}{unpack('din', 'd')}{prm}{_emit(self.second_ops)}{pack(self.second_out)}}}\n"
        )


def generate(psi, n_params: int, dim: int = 3, dofs_per_node: int | None = None, uses_values: bool = False) -> GeneratedLaw:
    """`psi(G, *params)` (or `psi(G, vals, *params)` with uses_values) on sympy symbols -> GeneratedLaw."""
    import sympy as sp

    dpn = dim if dofs_per_node is None else int(dofs_per_node)
    G = sp.Matrix(dpn, dim, lambda i, j: sp.Symbol(f"G{i}{j}", real=True))
    V = [sp.Symbol(f"V{c}", real=True) for c in range(dpn)]
    prm = [sp.Symbol(f"p{k}", real=True) for k in range(n_params)]
    expr = sp.sympify(psi(G, V, *prm) if uses_values else psi(G, *prm))
    in_names = [f"G{c}{j}" for c in range(dpn) for j in range(dim)] + ([f"V{c}" for c in range(dpn)] if uses_values else [])
    repl, (red,) = sp.cse([expr], optimizations="basic")

    def build():
        prog = _Prog()
        env = {sp.Symbol(n, real=True): n for n in in_names}
        env.update({p: str(p) for p in prm})
        for sym, sub in repl:
            env[sym] = _lower(sub, prog, env)
        return prog, _lower(red, prog, env)

    prog, out = build()
    psi_ops = _live(prog.ops, [out])
    prog1, out1 = build()
    grads = _reverse(prog1, out1, in_names)
    first_ops = _live(prog1.ops, list(grads.values()))
    prog2, out2 = build()
    grads2 = _reverse(prog2, out2, in_names)
    tan = _tangent(prog2, {n: "d" + n for n in in_names})
    second_out = {n: (ZERO if _is_const(g) else tan.get(g, ZERO)) for n, g in grads2.items()}
    second_ops = _live(prog2.ops, list(second_out.values()))
    return GeneratedLaw(dim, dpn, n_params, uses_values, psi_ops, first_ops, second_ops, out, grads, second_out)

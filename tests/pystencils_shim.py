"""tests/pystencils_shim.py -- TEST INFRASTRUCTURE (golden-vector generation only).

A stand-in for the five pystencils names the reference's ``pyst_kernels`` use, so that the
reference's OWN kernel definitions (``pyst_kernels/advection_flux.py:9-332``,
``advection_timestep.py:14-104``, ``elementwise_ops.py:9-104`` and the wrappers built on them)
can be imported unmodified and executed in a container where pystencils 1.0.1
(``poetry.lock:544-545``) cannot be installed.  ``tests/golden/make_golden_eno3.py`` installs this
module as ``sys.modules["pystencils"]`` before importing the reference.

What pystencils does with these kernels, and what the shim does instead:

* ``@ps.kernel`` re-parses the decorated function, turns every ``lhs @= rhs`` into an assignment
  and every ``a if c else b`` into a ``Piecewise`` and returns the assignment list.  The shim does
  the same AST rewrite (``__ps_assign__`` / ``__ps_select__``) and runs the function body with
  lazy expression nodes, so the assignment list is an expression tree in the **source's own
  evaluation order** (``(1 / 3) * field[0, 1] * velocity_x[0, 1]`` is ``((1/3) * f) * v``; the
  literals are Python doubles).  pystencils hands the tree to sympy and a C compiler with
  fast-math, which may re-associate -- the reference's result is therefore defined only up to
  rounding, and consumers of the goldens compare at 1e-13, not bitwise.
* ``ps.fields("a, b : float64[2D]")`` / ``[n0, n1]`` / ``[3D]`` / ``[2, n0, n1]`` -> field objects
  whose ``[offsets]`` are relative accesses.  Fixed shapes are checked at call time and a mismatch
  raises ``ValueError`` like the compiled pystencils kernel does.
* ``ps.create_kernel(assignments, config=...).compile()`` -> a callable taking keyword NumPy arrays
  and scalars.  Iteration space: pystencils' ``create_domain_kernel`` with ``ghost_layers=None``
  takes ``g = max(|offset|)`` over ALL field accesses of the kernel and strips ``g`` layers on both
  sides of EVERY axis (``make_loop_over_domain``; recalled from pystencils 1.0.x, it cannot be
  re-read offline -- SURVEY.md 8c "known oracle hazards").  The high side is forced anyway
  (``field[0, 2]`` is read); on the low side and on the transverse axis this is the rule the
  shim encodes, in one place: :func:`_ghost_layers`.
* ``ps.CreateKernelConfig(...)`` -> recorded, ignored (``cpu_openmp`` only changes threading).

Evaluation is whole-array NumPy over shifted views; every kernel here writes one field at offset
0 and reads the written field at offset 0 only, which the shim asserts, so whole-array evaluation
equals the per-cell loop.
"""
from __future__ import annotations

import ast
import inspect
import re
import textwrap

import numpy as np

__version__ = "shim-1.0.1"


# ---------------------------------------------------------------------------------------------
# expression nodes
# ---------------------------------------------------------------------------------------------
class Expr:
    __slots__ = ("op", "args")

    def __init__(self, op, *args):
        self.op, self.args = op, args

    # arithmetic, mirroring Python's own evaluation order
    def __add__(self, o): return Expr("add", self, o)
    def __radd__(self, o): return Expr("add", o, self)
    def __sub__(self, o): return Expr("sub", self, o)
    def __rsub__(self, o): return Expr("sub", o, self)
    def __mul__(self, o): return Expr("mul", self, o)
    def __rmul__(self, o): return Expr("mul", o, self)
    def __truediv__(self, o): return Expr("div", self, o)
    def __rtruediv__(self, o): return Expr("div", o, self)
    def __neg__(self): return Expr("neg", self)
    def __pos__(self): return self
    def __gt__(self, o): return Expr("gt", self, o)
    def __lt__(self, o): return Expr("lt", self, o)
    def __ge__(self, o): return Expr("ge", self, o)
    def __le__(self, o): return Expr("le", self, o)

    def __bool__(self):
        raise TypeError("a pystencils expression has no truth value; `a if c else b` must be rewritten")


class Access(Expr):
    __slots__ = ("field", "offsets")

    def __init__(self, field, offsets):
        Expr.__init__(self, "access")
        self.field, self.offsets = field, tuple(int(o) for o in offsets)


class Field:
    def __init__(self, name, dtype, ndim, fixed_shape):
        self.name, self.dtype, self.ndim, self.fixed_shape = name, dtype, ndim, fixed_shape

    def __getitem__(self, offsets):
        if not isinstance(offsets, tuple):
            offsets = (offsets,)
        assert len(offsets) == self.ndim, f"{self.name}: {len(offsets)} offsets for a {self.ndim}-D field"
        return Access(self, offsets)


def fields(description):
    """``"a, b : float64[2D]"``, ``"f : float64[128, 256]"``, ``"v : float64[2, 8, 8]"``."""
    m = re.fullmatch(r"\s*([^:]+?)\s*:\s*(\w+)\s*\[([^\]]+)\]\s*", description)
    assert m, f"unsupported field description {description!r}"
    names = [n.strip() for n in m.group(1).split(",")]
    dtype, shape = m.group(2), m.group(3).strip()
    md = re.fullmatch(r"(\d)D", shape)
    if md:
        ndim, fixed = int(md.group(1)), None
    else:
        fixed = tuple(int(s) for s in shape.split(","))
        ndim = len(fixed)
    out = tuple(Field(n, dtype, ndim, fixed) for n in names)
    return out[0] if len(out) == 1 else out


# ---------------------------------------------------------------------------------------------
# @ps.kernel
# ---------------------------------------------------------------------------------------------
class _Rewrite(ast.NodeTransformer):
    def visit_AugAssign(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.MatMult):
            target = ast.fix_missing_locations(ast.copy_location(
                ast.Subscript(value=node.target.value, slice=node.target.slice, ctx=ast.Load()), node.target))
            call = ast.Call(func=ast.Name(id="__ps_assign__", ctx=ast.Load()), args=[target, node.value], keywords=[])
            return ast.copy_location(ast.Expr(value=call), node)
        return node

    def visit_IfExp(self, node):
        self.generic_visit(node)
        return ast.copy_location(
            ast.Call(func=ast.Name(id="__ps_select__", ctx=ast.Load()), args=[node.test, node.body, node.orelse],
                     keywords=[]), node)


def kernel(func):
    """The decorator: returns the list of ``(lhs_access, rhs_expr)`` the function body states."""
    src = textwrap.dedent(inspect.getsource(func))
    tree = ast.parse(src)
    fdef = tree.body[0]
    assert isinstance(fdef, ast.FunctionDef)
    fdef.decorator_list = []
    tree = ast.fix_missing_locations(_Rewrite().visit(tree))
    assignments = []
    ns = dict(func.__globals__)
    cv = inspect.getclosurevars(func)
    ns.update(cv.nonlocals)
    ns["__ps_assign__"] = lambda lhs, rhs: assignments.append((lhs, rhs))
    ns["__ps_select__"] = lambda c, a, b: Expr("select", c, a, b)
    exec(compile(tree, inspect.getsourcefile(func) or "<ps.kernel>", "exec"), ns)
    ns[fdef.name]()
    assert assignments, f"{fdef.name}: no `@=` assignment found"
    return assignments


# ---------------------------------------------------------------------------------------------
# create_kernel(...).compile()
# ---------------------------------------------------------------------------------------------
class CreateKernelConfig:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _accesses(e, out):
    if isinstance(e, Access):
        out.append(e)
    elif isinstance(e, Expr):
        for a in e.args:
            _accesses(a, out)
    return out


def _ghost_layers(accesses):
    """pystencils: ``max(fa.required_ghost_layers)``, ``required_ghost_layers = max(|offsets|)``; the
    same count is stripped on both sides of every axis."""
    return max((max((abs(o) for o in a.offsets), default=0) for a in accesses), default=0)


class _Compiled:
    def __init__(self, assignments, config):
        self.assignments, self.config = assignments, config
        acc = []
        for lhs, rhs in assignments:
            _accesses(lhs, acc)
            _accesses(rhs, acc)
        self.ghost = _ghost_layers(acc)
        self.fields = {a.field.name: a.field for a in acc}
        written = {lhs.field.name for lhs, _ in assignments}
        for a in acc:
            if a.field.name in written:
                assert all(o == 0 for o in a.offsets), "shim: a written field may only be accessed at offset 0"

    def __call__(self, **kw):
        arrays = {}
        shape = None
        for name, f in self.fields.items():
            if name not in kw:
                raise KeyError(f"Mismatch of field names, expected {sorted(self.fields)}")
            a = kw[name]
            if not isinstance(a, np.ndarray) or a.ndim != f.ndim:
                raise ValueError(f"Wrong number of dimensions for argument {name}")
            if f.fixed_shape is not None and tuple(a.shape) != f.fixed_shape:
                raise ValueError(f"Wrong shape for array {name}: expected {f.fixed_shape}, got {a.shape}")
            if a.dtype != np.dtype(f.dtype):
                raise ValueError(f"Wrong data type for array {name}: expected {f.dtype}, got {a.dtype}")
            if shape is None:
                shape = a.shape
            elif a.shape != shape:
                raise ValueError(f"Wrong shape for array {name}: fields of one kernel share a shape")
            arrays[name] = a
        g = self.ghost
        if any(n <= 2 * g for n in shape):
            return

        def view(acc):
            a = arrays[acc.field.name]
            return a[tuple(slice(g + o, n - g + o) for o, n in zip(acc.offsets, a.shape))]

        def ev(e):
            if isinstance(e, Access):
                return view(e)
            if isinstance(e, Expr):
                op = e.op
                if op == "select":
                    c, a, b = (ev(x) for x in e.args)
                    return np.where(c, a, b)
                if op == "neg":
                    return -ev(e.args[0])
                x, y = ev(e.args[0]), ev(e.args[1])
                return {"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.true_divide,
                        "gt": np.greater, "lt": np.less, "ge": np.greater_equal, "le": np.less_equal}[op](x, y)
            if isinstance(e, (int, float, np.floating, np.integer)):
                return e
            # a sympy symbol (or a sympy expression of symbols, e.g. -inv_dx): value from the keywords
            syms = sorted(e.free_symbols, key=str)
            vals = {}
            for s in syms:
                if str(s) not in kw:
                    raise KeyError(f"missing scalar argument {s}")
                vals[s] = kw[str(s)]
            if len(syms) == 1 and e == syms[0]:
                return vals[syms[0]]
            import sympy as sp

            return sp.lambdify(syms, e, "math")(*[vals[s] for s in syms])

        for lhs, rhs in self.assignments:
            val = ev(rhs)
            view(lhs)[...] = val


class _Ast:
    def __init__(self, assignments, config):
        self.assignments, self.config = assignments, config

    def compile(self):
        return _Compiled(self.assignments, self.config)


def create_kernel(assignments, config=None, **kw):
    return _Ast(assignments, config)


def install():
    """Register this module as ``pystencils`` (only if the real one is absent)."""
    import sys

    try:
        import pystencils  # noqa: F401
        return False
    except ImportError:
        sys.modules["pystencils"] = sys.modules[__name__]
        return True

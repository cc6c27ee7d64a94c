"""Translate a reference-style Python right-hand side into a CUDA C++ snippet.

The reference takes numba-compiled functions `f(t, y) -> dy` (`_aux.py:64-65`) or, through
`Advanced`, `f(t, y, parameters) -> (dy, auxiliary)` (`_advanced.py:22-26`).  A Python callable
cannot run inside a CUDA kernel, so the facade translates the function's *source* (its AST) into the
body of the device function that NVRTC fuses into the stepper (`ni.CudaRHS`).  Nothing is evaluated
on the host: the Python function is never called.

Supported subset (enough for the reference's README / test problems and typical small ODE systems):
scalar assignments, `np.array((e0, e1, ...))`, element-wise arithmetic on vectors and scalars
(`+ - * / **` with broadcasting of scalars), `y`, `y[i]`, `y.copy()`, `parameters[k]` for tuple or
array parameters, `np.zeros(k)/np.empty(k)` + `v[i] = ...`, `for i in range(const)` (unrolled),
`if/else` on scalar conditions, `math.` / `np.` elementary functions, names bound to numeric
constants in the function's globals / closure.  Anything else raises `TranslationError` naming the
construct.  Expressions are emitted fully parenthesised in source order and compiled with
`--fmad=false`, so the arithmetic is the unfused arithmetic numba generates.
"""
from __future__ import annotations

import ast
import inspect
import math
import textwrap

import numpy as np

_FUNCS = {'sin': 'sin', 'cos': 'cos', 'tan': 'tan', 'exp': 'exp', 'log': 'log', 'sqrt': 'sqrt', 'abs': 'fabs',
          'fabs': 'fabs', 'absolute': 'fabs', 'tanh': 'tanh', 'sinh': 'sinh', 'cosh': 'cosh', 'arctan': 'atan',
          'atan': 'atan', 'arcsin': 'asin', 'asin': 'asin', 'arccos': 'acos', 'acos': 'acos', 'log10': 'log10',
          'log2': 'log2', 'exp2': 'exp2', 'floor': 'floor', 'ceil': 'ceil', 'sign': 'ni_sign'}
_FUNCS2 = {'arctan2': 'atan2', 'atan2': 'atan2', 'power': 'pow', 'pow': 'pow', 'maximum': 'fmax', 'minimum': 'fmin',
           'hypot': 'hypot', 'fmod': 'fmod', 'copysign': 'copysign'}
_BINOPS = {ast.Add: '+', ast.Sub: '-', ast.Mult: '*', ast.Div: '/'}
_CMPOPS = {ast.Lt: '<', ast.LtE: '<=', ast.Gt: '>', ast.GtE: '>=', ast.Eq: '==', ast.NotEq: '!='}


class TranslationError(TypeError):
    pass


class Vec(list):
    """A vector value: list of C scalar expressions."""


def _lit(x) -> str:
    x = float(x)
    if math.isinf(x):
        return '(1.0/0.0)' if x > 0 else '(-1.0/0.0)'
    if math.isnan(x):
        return '(0.0/0.0)'
    r = repr(x)
    if 'e' not in r and '.' not in r and 'inf' not in r:
        r += '.0'
    return f'({r})' if x < 0 else r


class _Translator:
    def __init__(self, fun, n, parameters, advanced):
        self.pyfun = getattr(fun, 'py_func', fun)
        try:
            src = textwrap.dedent(inspect.getsource(self.pyfun))
        except (OSError, TypeError) as e:
            raise TranslationError(f'cannot read the source of {fun!r}: {e}') from e
        self.fdef = self._find_function(src)
        args = [a.arg for a in self.fdef.args.args]
        if len(args) not in (2, 3):
            raise TranslationError(f'the right-hand side must take (t, y) or (t, y, parameters), got {args}')
        # `f(t, y, parameters)` returns (dy, auxiliary) (_advanced.py:22-26); `f(t, y)` returns dy, whichever step
        # semantics the solver uses (ni.sr applies the Advanced semantics to a basic right-hand side)
        self.n, self.advanced = n, len(args) == 3
        self.tname, self.yname = args[0], args[1]
        self.pname = args[2] if len(args) == 3 else None
        self.env = {}          # python name -> C scalar expr | Vec | int constant | _Params
        self.decl = set()
        self.lines = []
        self.tmp = 0
        self.env[self.tname] = 't'
        self.env[self.yname] = Vec(f'y[{i}]' for i in range(n))
        self.P = 0
        if self.pname is not None:
            self.env[self.pname] = self._param_struct(parameters)
        self.globals = dict(getattr(self.pyfun, '__globals__', {}))
        if getattr(self.pyfun, '__closure__', None):
            for name, cell in zip(self.pyfun.__code__.co_freevars, self.pyfun.__closure__):
                try:
                    self.globals[name] = cell.cell_contents
                except ValueError:
                    pass
        self.dy = self.aux = None
        self.aux_scalar = False

    def _find_function(self, src):
        """The FunctionDef / Lambda node of the right-hand side inside `src` (inspect.getsource gives the whole
        statement for a lambda, possibly cut in the middle of a call that spans several lines)."""
        want = self.pyfun.__code__.co_argcount
        trees = []
        try:
            trees.append(ast.parse(src))
        except SyntaxError:
            pos = src.find('lambda')
            while pos >= 0:  # grow-and-shrink: the longest prefix starting at this `lambda` that parses
                for end in range(len(src), pos + 6, -1):
                    try:
                        trees.append(ast.parse(src[pos:end].strip().rstrip(','), mode='eval'))
                        break
                    except SyntaxError:
                        continue
                pos = src.find('lambda', pos + 6)
        for tree in trees:
            for nd in ast.walk(tree):
                if isinstance(nd, (ast.FunctionDef, ast.Lambda)) and len(nd.args.args) == want:
                    return nd
        raise TranslationError(f'cannot find the definition of {self.pyfun.__name__} in its source')

    # ---- parameters: a flat float64 vector on the device; tuples of scalars / arrays are laid out in order
    def _param_struct(self, parameters):
        if parameters is None:
            raise TranslationError('the function takes parameters but none were given')
        if isinstance(parameters, np.ndarray):
            par = parameters
            m = par.shape[-1] if par.ndim else 1
            self.P = m
            return Vec(f'p[{i}]' for i in range(m))
        if isinstance(parameters, (tuple, list)):
            items, off = [], 0
            for q in parameters:
                a = np.asarray(q, dtype=np.float64)
                if a.ndim == 0:
                    items.append(f'p[{off}]')
                    off += 1
                else:
                    items.append(Vec(f'p[{off + i}]' for i in range(a.size)))
                    off += a.size
            self.P = off
            return tuple(items)
        self.P = 1
        return 'p[0]'

    # ---- expressions
    def err(self, node, what):
        raise TranslationError(f'{what} (line {getattr(node, "lineno", "?")} of {self.pyfun.__name__}): '
                               'write the right-hand side as a ni.CudaRHS snippet instead')

    def const_int(self, node):
        v = self.expr(node)
        if isinstance(v, int):
            return v
        self.err(node, 'expected a compile-time integer')

    def broadcast(self, f, *vals):
        vecs = [v for v in vals if isinstance(v, Vec)]
        if not vecs:
            return f(*[self.scalar(v) for v in vals])
        m = len(vecs[0])
        if any(len(v) != m for v in vecs):
            raise TranslationError('vector operands of different lengths')
        return Vec(f(*[(v[i] if isinstance(v, Vec) else self.scalar(v)) for v in vals]) for i in range(m))

    def scalar(self, v):
        if isinstance(v, bool):
            return '1.0' if v else '0.0'
        if isinstance(v, (int, float)):
            return _lit(v)
        if isinstance(v, str):
            return v
        raise TranslationError(f'expected a scalar expression, got {type(v).__name__}')

    def expr(self, node):
        if isinstance(node, ast.Constant):
            if isinstance(node.value, bool) or not isinstance(node.value, (int, float)):
                self.err(node, f'unsupported constant {node.value!r}')
            return node.value
        if isinstance(node, ast.Name):
            if node.id in self.env:
                return self.env[node.id]
            if node.id in self.globals and isinstance(self.globals[node.id], (int, float, np.floating, np.integer)):
                g = self.globals[node.id]
                return int(g) if isinstance(g, (int, np.integer)) and not isinstance(g, bool) else float(g)
            self.err(node, f'unknown name {node.id!r}')
        if isinstance(node, ast.UnaryOp):
            v = self.expr(node.operand)
            if isinstance(node.op, ast.USub):
                if isinstance(v, (int, float)):
                    return -v
                return self.broadcast(lambda a: f'(-{a})', v)
            if isinstance(node.op, ast.UAdd):
                return v
            self.err(node, 'unsupported unary operator')
        if isinstance(node, ast.BinOp):
            a, b = self.expr(node.left), self.expr(node.right)
            if isinstance(node.op, ast.Pow):
                return self.power(node, a, b)
            if isinstance(node.op, (ast.Mod, ast.FloorDiv)):
                if isinstance(a, (int, float)) and isinstance(b, (int, float)):  # index arithmetic on loop counters
                    return a % b if isinstance(node.op, ast.Mod) else a // b
                if isinstance(node.op, ast.Mod):  # Python's float %: result has the sign of the divisor
                    return self.broadcast(lambda x, y: f'({x} - floor({x} / {y}) * {y})', a, b)
                return self.broadcast(lambda x, y: f'floor({x} / {y})', a, b)
            if type(node.op) in _BINOPS:
                op = _BINOPS[type(node.op)]
                if isinstance(a, (int, float)) and isinstance(b, (int, float)):  # constant folding as Python does
                    return {'+': a + b, '-': a - b, '*': a * b, '/': a / b if b else float('nan')}[op]
                return self.broadcast(lambda x, y: f'({x} {op} {y})', a, b)
            self.err(node, f'unsupported operator {type(node.op).__name__}')
        if isinstance(node, ast.Subscript):
            base = self.expr(node.value)
            if isinstance(node.slice, ast.Slice):
                if not isinstance(base, Vec):
                    self.err(node, 'slicing a non-vector')
                lo = 0 if node.slice.lower is None else self.const_int(node.slice.lower)
                hi = len(base) if node.slice.upper is None else self.const_int(node.slice.upper)
                return Vec(base[lo:hi])
            i = self.const_int(node.slice)
            if isinstance(base, (Vec, tuple)):
                try:
                    return base[i]
                except IndexError:
                    self.err(node, f'index {i} out of range')
            self.err(node, 'indexing a scalar')
        if isinstance(node, ast.Call):
            return self.call(node)
        if isinstance(node, (ast.Tuple, ast.List)):
            return Vec(self.scalar(self.expr(e)) for e in node.elts)
        if isinstance(node, ast.IfExp):
            folded = self._fold_cond(node.test)  # conditions on unrolled loop indices are decided here
            if folded is not None:
                return self.expr(node.body if folded else node.orelse)
            c = self.cond(node.test)
            return self.broadcast(lambda x, y: f'({c} ? {x} : {y})', self.expr(node.body), self.expr(node.orelse))
        if isinstance(node, ast.Attribute):
            if isinstance(node.value, ast.Name) and node.value.id in ('np', 'numpy', 'math'):
                if node.attr == 'pi':
                    return math.pi
                if node.attr == 'e':
                    return math.e
            base = self.expr(node.value)
            if node.attr == 'size' and isinstance(base, Vec):
                return len(base)
            self.err(node, f'unsupported attribute .{node.attr}')
        self.err(node, f'unsupported expression {type(node).__name__}')

    def power(self, node, a, b):
        if isinstance(a, (int, float)) and isinstance(b, (int, float)):
            return a ** b
        if isinstance(b, int) and 0 <= b <= 64:  # numba lowers literal integer powers to exponentiation by squaring
            def f(x):
                if b == 0:
                    return '1.0'
                result, base, e = None, x, b
                while e:
                    if e & 1:
                        result = base if result is None else f'({result} * {base})'
                    e >>= 1
                    if e:
                        base = f'({base} * {base})'
                return result
            return self.broadcast(f, a)
        return self.broadcast(lambda x, y: f'pow({x}, {y})', a, b)

    def call(self, node):
        f = node.func
        name = f.attr if isinstance(f, ast.Attribute) else f.id if isinstance(f, ast.Name) else None
        if isinstance(f, ast.Attribute) and f.attr == 'copy' and not node.args:
            return self.expr(f.value)
        if name == 'array' or name == 'asarray':
            v = self.expr(node.args[0])
            return v if isinstance(v, Vec) else Vec([self.scalar(v)])
        if name in ('zeros', 'empty', 'ones', 'zeros_like', 'empty_like'):
            if name.endswith('_like'):
                m = len(self.expr(node.args[0]))
            else:
                m = self.const_int(node.args[0])
            return Vec(['1.0' if name == 'ones' else '0.0'] * m)
        if name == 'len':
            v = self.expr(node.args[0])
            if isinstance(v, (Vec, tuple)):
                return len(v)
        if name == 'float' or name == 'float64':
            return self.expr(node.args[0])
        if name == 'sum' and len(node.args) == 1:
            v = self.expr(node.args[0])
            if isinstance(v, Vec):
                out = v[0]
                for x in v[1:]:
                    out = f'({out} + {x})'
                return out
        if name == 'dot' and len(node.args) == 2:
            a, b = self.expr(node.args[0]), self.expr(node.args[1])
            if isinstance(a, Vec) and isinstance(b, Vec) and len(a) == len(b):
                out = f'({a[0]} * {b[0]})'
                for x, y in zip(a[1:], b[1:]):
                    out = f'({out} + ({x} * {y}))'
                return out
        if name in _FUNCS and len(node.args) == 1:
            c = _FUNCS[name]
            return self.broadcast(lambda x: f'{c}({x})', self.expr(node.args[0]))
        if name in _FUNCS2 and len(node.args) == 2:
            c = _FUNCS2[name]
            return self.broadcast(lambda x, y: f'{c}({x}, {y})', self.expr(node.args[0]), self.expr(node.args[1]))
        if name in ('max', 'min') and len(node.args) == 2:
            c = 'fmax' if name == 'max' else 'fmin'
            return self.broadcast(lambda x, y: f'{c}({x}, {y})', self.expr(node.args[0]), self.expr(node.args[1]))
        self.err(node, f'unsupported call {ast.unparse(node.func)}()')

    def cond(self, node):
        if isinstance(node, ast.Compare) and len(node.ops) == 1 and type(node.ops[0]) in _CMPOPS:
            a, b = self.scalar(self.expr(node.left)), self.scalar(self.expr(node.comparators[0]))
            return f'({a} {_CMPOPS[type(node.ops[0])]} {b})'
        if isinstance(node, ast.BoolOp):
            op = ' && ' if isinstance(node.op, ast.And) else ' || '
            return '(' + op.join(self.cond(v) for v in node.values) + ')'
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.Not):
            return f'(!{self.cond(node.operand)})'
        self.err(node, 'unsupported condition')

    # ---- statements
    def materialise(self, name, value):
        """Bind a Python name to C variables holding `value` (so later reassignment inside if/for works)."""
        if getattr(self, 'rt_depth', 0) > 0 and isinstance(self.env.get(name), (int, tuple)) \
                and not isinstance(self.env.get(name), bool):
            # translate-time constants are folded into the expressions that use them: rebinding one under a run-time
            # condition would silently take effect unconditionally
            raise TranslationError(f'`{name}` holds a translate-time constant (an int or a tuple) and is reassigned '
                                   f'inside a run-time `if`; write it as a float (e.g. `{name} = 2.0`) so that it '
                                   'becomes a variable of the generated code')
        if isinstance(value, (int,)) and not isinstance(value, bool):
            self.env[name] = value
            if getattr(self, 'rt_depth', 0) > 0:
                self.rt_scoped[-1].add(name)  # a constant bound under a run-time condition dies with its branch
            return
        if isinstance(value, tuple):
            self.env[name] = value
            if getattr(self, 'rt_depth', 0) > 0:
                self.rt_scoped[-1].add(name)
            return
        if isinstance(value, Vec):
            names = Vec(f'v_{name}_{i}' for i in range(len(value)))
            vals = [self.scalar(ce) for ce in value]
            if any(cn in v for cn in names for v in vals):  # e.g. v = v[::-1]-like self reference: go through temporaries
                tmps = [f'{cn}_n' for cn in names]
                for tn, v in zip(tmps, vals):
                    self.lines.append(f'{tn} = {v};')
                    self.decl.add(tn)
                vals = tmps
            for cn, v in zip(names, vals):
                self.lines.append(f'{cn} = {v};')
                self.decl.add(cn)
            self.env[name] = names
            return
        cn = f's_{name}'
        self.lines.append(f'{cn} = {self.scalar(value)};')
        self.decl.add(cn)
        self.env[name] = cn

    def stmt(self, node):
        if isinstance(node, ast.Expr) and isinstance(node.value, ast.Constant):
            return  # docstring
        if isinstance(node, ast.Pass):
            return
        if isinstance(node, ast.Assign) and len(node.targets) == 1:
            tgt = node.targets[0]
            if isinstance(tgt, ast.Name):
                self.materialise(tgt.id, self.expr(node.value))
                return
            if isinstance(tgt, ast.Subscript) and isinstance(tgt.value, ast.Name):
                base = self.env.get(tgt.value.id)
                i = self.const_int(tgt.slice)
                if isinstance(base, Vec):
                    self.lines.append(f'{base[i]} = {self.scalar(self.expr(node.value))};')
                    return
            if isinstance(tgt, ast.Tuple) and isinstance(node.value, ast.Tuple) and len(tgt.elts) == len(node.value.elts):
                vals = [self.expr(v) for v in node.value.elts]
                for t_, v in zip(tgt.elts, vals):
                    if not isinstance(t_, ast.Name):
                        self.err(node, 'unsupported assignment target')
                    self.materialise(t_.id, v)
                return
            if isinstance(tgt, ast.Tuple):
                v = self.expr(node.value)
                if isinstance(v, (Vec, tuple)) and len(v) == len(tgt.elts):
                    for t_, x in zip(tgt.elts, v):
                        self.materialise(t_.id, x)
                    return
            self.err(node, 'unsupported assignment target')
        if isinstance(node, ast.AugAssign) and isinstance(node.target, ast.Name) and type(node.op) in _BINOPS:
            cur = self.expr(node.target)
            op = _BINOPS[type(node.op)]
            new = self.broadcast(lambda x, y: f'({x} {op} {y})', cur, self.expr(node.value))
            self.materialise(node.target.id, new)
            return
        if isinstance(node, ast.AugAssign) and isinstance(node.target, ast.Subscript) and type(node.op) in _BINOPS:
            base = self.expr(node.target.value)
            i = self.const_int(node.target.slice)
            op = _BINOPS[type(node.op)]
            if isinstance(base, Vec):
                self.lines.append(f'{base[i]} = ({base[i]} {op} {self.scalar(self.expr(node.value))});')
                return
        if isinstance(node, ast.For) and isinstance(node.target, ast.Name) and isinstance(node.iter, ast.Call) \
                and getattr(node.iter.func, 'id', None) == 'range' and not node.orelse:
            args = [self.const_int(a) for a in node.iter.args]
            rng = range(*args)
            if len(rng) > 4096:
                self.err(node, 'loop too long to unroll')
            for i in rng:
                self.env[node.target.id] = i
                for st in node.body:
                    self.stmt(st)
            return
        if isinstance(node, ast.If):
            test = node.test
            # compile-time conditions on unrolled loop indices
            folded = self._fold_cond(test)
            if folded is not None:
                for st in (node.body if folded else node.orelse):
                    self.stmt(st)
                return
            c = self.cond(test)
            self.lines.append(f'if {c} {{')
            self._runtime_branch(node.body)
            if node.orelse:
                self.lines.append('} else {')
                self._runtime_branch(node.orelse)
            self.lines.append('}')
            return
        if isinstance(node, ast.Return):
            if getattr(self, 'rt_depth', 0) > 0:
                # the outputs are evaluated once, after all statements: a return under a run-time condition would be
                # hoisted out of it
                self.err(node, '`return` inside a run-time `if` (select the value into a variable and return once)')
            self.ret(node)
            return
        self.err(node, f'unsupported statement {type(node).__name__}')

    def _runtime_branch(self, body):
        """Statements of one branch of a run-time `if`: translate-time constants first bound here are dropped again."""
        self.rt_depth = getattr(self, 'rt_depth', 0) + 1
        if not hasattr(self, 'rt_scoped'):
            self.rt_scoped = []
        self.rt_scoped.append(set())
        try:
            for st in body:
                self.stmt(st)
        finally:
            for name in self.rt_scoped.pop():
                self.env.pop(name, None)
            self.rt_depth -= 1

    def _fold_cond(self, test):
        """True / False if the condition only involves compile-time numbers, else None."""
        if isinstance(test, ast.Compare) and len(test.ops) == 1 and type(test.ops[0]) in _CMPOPS:
            try:
                a, b = self.expr(test.left), self.expr(test.comparators[0])
            except TranslationError:
                return None
            if isinstance(a, (int, float)) and isinstance(b, (int, float)):
                import operator
                return {'<': operator.lt, '<=': operator.le, '>': operator.gt, '>=': operator.ge, '==': operator.eq,
                        '!=': operator.ne}[_CMPOPS[type(test.ops[0])]](a, b)
        return None

    def ret(self, node):
        if self.dy is not None:
            self.err(node, 'more than one return statement')
        v = node.value
        if self.advanced:
            if not (isinstance(v, ast.Tuple) and len(v.elts) == 2):
                self.err(node, 'an Advanced right-hand side must `return dy, auxiliary`')
            dy, aux = self.expr(v.elts[0]), v.elts[1]
            if isinstance(aux, ast.Tuple):
                parts = [self.expr(e) for e in aux.elts]
                flat = Vec()
                for q in parts:
                    flat.extend(q if isinstance(q, Vec) else [self.scalar(q)])
                self.aux = flat
            else:
                a = self.expr(aux)
                if isinstance(a, tuple):  # e.g. `return dy, p` with tuple parameters
                    flat = Vec()
                    for q in a:
                        flat.extend(q if isinstance(q, Vec) else [self.scalar(q)])
                    self.aux = flat
                elif isinstance(a, Vec):
                    self.aux = a
                else:
                    self.aux = Vec([self.scalar(a)])
                    self.aux_scalar = True
        else:
            dy = self.expr(v)
            self.aux = Vec()
        if not isinstance(dy, Vec):
            dy = Vec([self.scalar(dy)])
        if len(dy) != self.n:
            raise TranslationError(f'the function returns {len(dy)} derivatives for {self.n} state components')
        self.dy = dy

    def run(self):
        if isinstance(self.fdef, ast.Lambda):
            self.ret(ast.Return(value=self.fdef.body))
        else:
            for st in self.fdef.body:
                self.stmt(st)
        if self.dy is None:
            raise TranslationError('the function has no return statement')
        # evaluate every output into a temporary first: outputs may alias inputs (dy = y)
        out = [f'double {cn} = 0.0;' for cn in sorted(self.decl)] + list(self.lines)  # declarations hoisted: function scope
        for i, e in enumerate(self.dy):
            out.append(f'const double o_dy{i} = {self.scalar(e)};')
        for i, e in enumerate(self.aux):
            out.append(f'const double o_aux{i} = {self.scalar(e)};')
        out += [f'dy[{i}] = o_dy{i};' for i in range(len(self.dy))]
        out += [f'aux[{i}] = o_aux{i};' for i in range(len(self.aux))]
        return '\n'.join('    ' + ln for ln in out)


def translate(fun, n, parameters=None, advanced=False):
    """Returns (cuda_body, P, Q, aux_is_scalar) for a reference-style right-hand side `fun`."""
    tr = _Translator(fun, int(n), parameters, bool(advanced))
    body = tr.run()
    return body, tr.P, len(tr.aux), tr.aux_scalar


class _CondTranslator(_Translator):
    """`condition(solver, parameters) -> bool` (_API.py:41-49): `solver.t`, `solver.y`, `solver.auxiliary` and the
    predicate parameters are the only inputs a device predicate can see."""

    def __init__(self, fun, n, Q, aux_scalar, parameters):
        self.pyfun = getattr(fun, 'py_func', fun)
        try:
            src = textwrap.dedent(inspect.getsource(self.pyfun))
        except (OSError, TypeError) as e:
            raise TranslationError(f'cannot read the source of {fun!r}: {e}') from e
        self.fdef = self._find_function(src)
        args = [a.arg for a in self.fdef.args.args]
        if len(args) != 2:
            raise TranslationError(f'a predicate takes (solver, parameters), got {args}')
        self.n, self.advanced = n, False
        self.sname, self.pname = args
        self.env, self.decl, self.lines = {}, set(), []
        par = np.atleast_1d(np.asarray(parameters, dtype=np.float64)) if not isinstance(parameters, (tuple, list)) \
            else np.concatenate([np.ravel(np.asarray(q, dtype=np.float64)) for q in parameters])
        if par.size > 8:
            raise TranslationError('a device predicate takes at most 8 parameters')
        self.cond_params = par
        if isinstance(parameters, (tuple, list)) or np.ndim(parameters) > 0:
            self.env[self.pname] = Vec(f'cp[{i}]' for i in range(par.size))
        else:
            self.env[self.pname] = 'cp[0]'
        aux = Vec(f'aux[{i}]' for i in range(Q))
        self.attrs = {'t': 't', 'y': Vec(f'y[{i}]' for i in range(n)),
                      'auxiliary': (aux[0] if aux_scalar and Q == 1 else aux)}
        self.globals = dict(getattr(self.pyfun, '__globals__', {}))
        self.result = None

    def expr(self, node):
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == self.sname:
            if node.attr in self.attrs:
                return self.attrs[node.attr]
            self.err(node, f'a device predicate can read solver.t, solver.y and solver.auxiliary, not solver.{node.attr}')
        return super().expr(node)

    def ret(self, node):
        if self.result is not None:
            self.err(node, 'more than one return statement')
        v = node.value
        self.result = self.cond(v) if isinstance(v, (ast.Compare, ast.BoolOp)) or \
            (isinstance(v, ast.UnaryOp) and isinstance(v.op, ast.Not)) else f'({self.scalar(self.expr(v))} != 0.0)'

    def run(self):
        if isinstance(self.fdef, ast.Lambda):
            self.ret(ast.Return(value=self.fdef.body))
        else:
            for st in self.fdef.body:
                self.stmt(st)
        if self.result is None:
            raise TranslationError('the predicate has no return statement')
        out = [f'double {cn} = 0.0;' for cn in sorted(self.decl)] + list(self.lines) + [f'return {self.result};']
        return '\n'.join('    ' + ln for ln in out)


def translate_condition(fun, n, Q=0, aux_scalar=False, parameters=0.0):
    """Returns (cuda_body, parameter_vector) for a reference-style predicate `fun(solver, parameters)`."""
    tr = _CondTranslator(fun, int(n), int(Q), bool(aux_scalar), parameters)
    return tr.run(), tr.cond_params

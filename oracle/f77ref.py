"""Runs the REFERENCE'S OWN Fortran source of the hot path, unmodified, through a small fixed-form
Fortran-77/90 subset translator (this file) -- TEST INFRASTRUCTURE ONLY.

There is no Fortran compiler in this image, so `oracle/_ref` cannot be a compiled binary.  Instead the
reference's source text (`/root/reference/pom/solver.f`, `advance.f`, `bounds_forcing.f` with the COMMON
declarations of `/root/reference/pom.h_dist`) is read where it lies, translated statement by statement to
Python and executed: every formula, loop range and branch is the reference's, nothing is restated by hand.
The translation keeps the semantics of the reference's build (`makefile_dist:17`, gfortran -O0):
  * `double precision` = IEEE binary64 (numpy float64), strict left-to-right evaluation, no FMA contraction;
  * literals without a `d` exponent are SINGLE precision (numpy float32) and follow Fortran's promotion
    rules (`(2./3.)` is evaluated in float32; `1./24.`, `(15.8*cbcnst)**(2./3.)`);
  * integer / integer truncates; `x**n` with an integer n is repeated multiplication (gfortran's powi
    expansion), real powers go through libm `pow`; `real(x,16)` is numpy's long double;
  * automatic (stack) arrays start at zero; whole-array statements and array sections are numpy slices.
`exchange2d_mpi` / `exchange3d_mpi` / `order*_mpi` are no-ops (single sub-domain: every neighbour is -1,
parallel_mpi.f:154-351), print / netCDF routines are not on the path.

It is slow (pure Python loops): use it on grids of a few hundred columns.  `scripts/make_ref_golden.py`
generates the fixtures under tests/golden/ref_*.npz with it; tests compare the C oracle and the CUDA path
with those fixtures, and -- when /root/reference is present -- with a live run.
"""
import math
import os
import re

import numpy as np

REF_ROOT = os.environ.get("POM_REFERENCE", "/root/reference")
LIVE_GRIDS = ((13, 11, 6), (16, 14, 7))     # grids of the live reference runs in tests/ (built into oracle/_ref by build())

f8 = np.float64
f4 = np.float32
f16 = np.longdouble


# ------------------------------------------------------------------------------------------------
# source reader: fixed form -> logical statements
def read_statements(path):
    out = []
    for ln, raw in enumerate(open(path, errors="replace"), 1):
        line = raw.rstrip("\n").expandtabs(8)
        if not line.strip():
            continue
        if line[0] in "cC*!":
            continue
        # strip inline comment (no character literals with '!' on the path)
        body = line
        q = None
        for p, ch in enumerate(line):
            if ch in "'\"":
                q = None if q == ch else (ch if q is None else q)
            elif ch == "!" and q is None:
                body = line[:p]
                break
        if not body.strip():
            continue
        if len(body) > 5 and body[5] not in " 0" and body[:5].strip() == "":
            if not out:
                raise SyntaxError(f"{path}:{ln}: continuation without a statement")
            out[-1] = (out[-1][0], out[-1][1] + body[6:])
        else:
            out.append((ln, body[6:] if len(body) > 6 else ""))
    return [(ln, s.strip()) for ln, s in out if s.strip()]


# ------------------------------------------------------------------------------------------------
# expression parser
TOK = re.compile(r"""\s*(?:
    (?P<num>(?:\d+\.(?!(?:lt|le|gt|ge|eq|ne|and|or|not|eqv|neqv)\.)\d*|\.\d+|\d+)(?:[de][+-]?\d+)?(?:_\w+)?)
  | (?P<dotop>\.(?:lt|le|gt|ge|eq|ne|and|or|not|true|false|eqv|neqv)\.)
  | (?P<id>[a-z_][a-z0-9_]*)
  | (?P<str>'[^']*'|"[^"]*")
  | (?P<op>\*\*|==|/=|<=|>=|::|[-+*/(),:<>=])
)""", re.X)


def tokenize(s):
    s = s.lower()
    pos, toks = 0, []
    while pos < len(s):
        if s[pos:].strip() == "":
            break
        m = TOK.match(s, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize {s[pos:pos + 20]!r} in {s!r}")
        pos = m.end()
        kind = m.lastgroup
        text = m.group(kind)
        # "1.and." style ambiguity does not occur on the path; but "2.d0" is handled by `num`
        toks.append((kind, text))
    return toks


class P:   # precedence-climbing parser producing a small AST (tuples)
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def expect(self, text):
        k, t = self.next()
        if t != text:
            raise SyntaxError(f"expected {text!r}, got {t!r} in {self.t}")

    def expr(self):
        return self.p_or()

    def p_or(self):
        a = self.p_and()
        while self.peek()[1] == ".or.":
            self.next(); a = ("or", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek()[1] == ".and.":
            self.next(); a = ("and", a, self.p_not())
        return a

    def p_not(self):
        if self.peek()[1] == ".not.":
            self.next(); return ("not", self.p_not())
        return self.p_rel()

    REL = {".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".eq.": "==", ".ne.": "!=",
           "<": "<", "<=": "<=", ">": ">", ">=": ">=", "==": "==", "/=": "!="}

    def p_rel(self):
        a = self.p_add()
        if self.peek()[1] in self.REL:
            op = self.REL[self.next()[1]]
            return ("rel", op, a, self.p_add())
        return a

    def p_add(self):
        if self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            a = self.p_mul()
            a = ("neg", a) if op == "-" else a
        else:
            a = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            a = ("bin", op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            a = ("bin", op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_atom()
        if self.peek()[1] == "**":
            self.next()
            # right associative; the exponent may carry a sign
            if self.peek()[1] in ("+", "-"):
                op = self.next()[1]
                b = self.p_pow()
                b = ("neg", b) if op == "-" else b
            else:
                b = self.p_pow()
            return ("pow", a, b)
        return a

    def p_atom(self):
        k, t = self.next()
        if k == "num":
            return ("num", t)
        if k == "dotop" and t in (".true.", ".false."):
            return ("bool", t == ".true.")
        if k == "str":
            return ("str", t[1:-1])
        if t == "(":
            e = self.expr()
            self.expect(")")
            return ("par", e)
        if k == "id":
            if self.peek()[1] == "(":
                self.next()
                args = []
                if self.peek()[1] != ")":
                    while True:
                        args.append(self.arg())
                        if self.peek()[1] == ",":
                            self.next(); continue
                        break
                self.expect(")")
                return ("call", t, args)
            return ("id", t)
        raise SyntaxError(f"unexpected token {t!r} in {self.t}")

    def arg(self):   # expression, or a section a:b / : / a: / :b
        lo = hi = None
        if self.peek()[1] == ":":
            self.next()
            if self.peek()[1] not in (",", ")"):
                hi = self.expr()
            return ("sec", None, hi)
        lo = self.expr()
        if self.peek()[1] == ":":
            self.next()
            if self.peek()[1] not in (",", ")"):
                hi = self.expr()
            return ("sec", lo, hi)
        return lo


def parse_expr(s):
    p = P(tokenize(s))
    e = p.expr()
    if p.i != len(p.t):
        raise SyntaxError(f"trailing tokens in {s!r}")
    return e


# ------------------------------------------------------------------------------------------------
# runtime helpers of the generated code
def _isint(x):
    return isinstance(x, (int, np.integer)) and not isinstance(x, bool)


def _div(a, b):
    if _isint(a) and _isint(b):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    return a / b


def _powi(x, n):
    # gfortran expands x**n (integer n) into multiplications by repeated squaring (__builtin_powi)
    if n < 0:
        return type(x)(1) / _powi(x, -n) if not _isint(x) else 0
    r, first = None, True
    y = x
    while n:
        if n & 1:
            r = y if r is None else r * y
        n >>= 1
        if n:
            y = y * y
    return (type(x)(1) if not _isint(x) else 1) if r is None else r


def _pow(a, b):
    if _isint(b):
        return _powi(a, int(b))
    if isinstance(a, np.ndarray):
        return np.power(a, b)
    if isinstance(a, f4) and isinstance(b, f4):
        return f4(math.pow(float(a), float(b)))       # powf: correctly rounded from the double result here
    return f8(math.pow(float(a), float(b)))


def _sign(a, b):
    r = abs(a)
    return r if b >= 0 else -r


def _max(*a):
    r = a[0]
    for x in a[1:]:
        r = x if x > r else r
    return r if len({type(x) for x in a}) == 1 else f8(r)


def _min(*a):
    r = a[0]
    for x in a[1:]:
        r = x if x < r else r
    return r if len({type(x) for x in a}) == 1 else f8(r)


def _real(x, kind=None):
    if kind is None or kind == 4:
        return f4(x)
    if kind == 8:
        return f8(x)
    return f16(x)


def _exp(x):
    if isinstance(x, np.longdouble):
        return np.exp(x)
    if isinstance(x, f4):
        return f4(math.exp(float(x)))
    return f8(math.exp(float(x)))


def _sqrt(x):
    return np.sqrt(x)


def _abs(x):
    return abs(x)


def _log(x):
    return f8(math.log(float(x)))


def _sin(x):      # initialize.f:349 (Coriolis parameter); not on the hot path
    return f8(math.sin(float(x)))


INTRINSICS = {"abs": "_abs", "dabs": "_abs", "sqrt": "_sqrt", "dsqrt": "_sqrt", "max": "_max", "min": "_min",
              "dmax1": "_max", "dmin1": "_min", "amax1": "_max", "amin1": "_min", "exp": "_exp", "sign": "_sign",
              "mod": "_mod", "float": "f4", "dble": "f8", "real": "_real", "int": "int", "log": "_log",
              "maxval": "np.max", "minval": "np.min", "sum": "_sum", "nint": "_nint", "sin": "_sin"}


def _sum(a):
    # gfortran's SUM: one accumulator, array element order (numpy's sum is pairwise)
    r = f8(0.)
    for x in np.asarray(a).ravel(order="F"):
        r = r + x
    return r


def _nint(x):
    return int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


def _mod(a, b):
    if _isint(a) and _isint(b):
        return int(math.fmod(a, b))
    return math.fmod(a, b)


RUNTIME = dict(np=np, f4=f4, f8=f8, f16=f16, _div=_div, _pow=_pow, _sign=_sign, _max=_max, _min=_min, _real=_real,
               _exp=_exp, _sqrt=_sqrt, _abs=_abs, _log=_log, _mod=_mod, _nint=_nint, _sum=_sum, int=int, _sin=_sin)


# ------------------------------------------------------------------------------------------------
# declarations
TYPES = (("double precision", "f8"), ("integer", "i"), ("real", "f4"), ("logical", "b"), ("character", "c"))


def split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def parse_decl(stmt):
    """-> (type, [(name, dims or None)]) or None.  dims = list of (lo_expr or None, hi_expr) strings."""
    s = stmt.lower()
    for kw, ty in TYPES:
        if s.startswith(kw):
            rest = s[len(kw):]
            if kw == "character" or kw == "real":
                m = re.match(r"\s*(\*\s*\d+|\([^)]*\))", rest)   # character*26, character(len=256), real(kind=..)
                if m:
                    if kw == "real" and "16" in m.group(1):
                        ty = "f16"
                    elif kw == "real" and "8" in m.group(1):
                        ty = "f8"
                    rest = rest[m.end():]
            if rest[:1] not in (" ", ",", ":") and rest[:1].isalnum():
                return None                               # e.g. an assignment to a variable called `integerx`
            common_dims = None
            if "::" in rest:
                attrs, rest = rest.split("::", 1)
                m = re.search(r"dimension\s*\((.*)\)", attrs)
                if m:
                    # strip anything after the matching parenthesis of dimension(
                    inner, depth = "", 0
                    for ch in attrs[attrs.index("dimension") + len("dimension"):].lstrip()[1:]:
                        if ch == "(":
                            depth += 1
                        if ch == ")":
                            if depth == 0:
                                break
                            depth -= 1
                        inner += ch
                    common_dims = inner
            names = []
            for item in split_top(rest):
                m = re.match(r"([a-z_][a-z0-9_]*)\s*(?:\((.*)\))?\s*(?:\*\s*\d+)?$", item)
                if not m:
                    raise SyntaxError(f"declaration item {item!r} in {stmt!r}")
                d = m.group(2) if m.group(2) is not None else common_dims
                dims = None
                if d is not None:
                    dims = []
                    for one in split_top(d):
                        if ":" in one:
                            lo, hi = one.split(":", 1)
                            dims.append((lo.strip(), hi.strip()))
                        else:
                            dims.append((None, one.strip()))
                names.append((m.group(1), dims))
            return ty, names
    return None


class Globals:
    """The COMMON state of pom.h for given grid parameters."""

    def __init__(self, header, params):
        self.types, self.dims, self.params = {}, {}, {}
        stmts = read_statements(header)
        for ln, s in stmts:
            low = s.lower()
            if low.startswith("parameter"):
                inner = low[low.index("(") + 1:low.rindex(")")]
                for item in split_top(inner):
                    k, v = item.split("=")
                    self.params[k.strip()] = v.strip()
            elif low.startswith("common"):
                m = re.match(r"common\s*/\s*\w+\s*/(.*)", low)
                for item in split_top(m.group(1)):
                    mm = re.match(r"([a-z_][a-z0-9_]*)\s*(?:\((.*)\))?$", item)
                    if mm.group(2) is not None:
                        self.dims[mm.group(1)] = [(None, x.strip()) if ":" not in x else tuple(y.strip() for y in x.split(":", 1))
                                                  for x in split_top(mm.group(2))]
                    self.dims.setdefault(mm.group(1), None)
            else:
                d = parse_decl(s)
                if d:
                    ty, names = d
                    for n, dm in names:
                        self.types[n] = ty
                        if dm is not None:
                            self.dims[n] = dm
        self.v = {}
        for k, v in params.items():
            self.v[k] = int(v)
        for k, expr in self.params.items():
            if k not in self.v:
                self.v[k] = int(eval(expr, {}, self.v))
        # allocate
        for n in self.dims:
            ty = self.types.get(n, "f8")
            if ty == "c":
                self.v[n] = ""
                continue
            dm = self.dims[n]
            if dm is None:
                self.v.setdefault(n, 0 if ty == "i" else (False if ty == "b" else f8(0.)))
            else:
                shape = tuple(int(eval(hi, {}, self.v)) - (int(eval(lo, {}, self.v)) if lo else 1) + 1 for lo, hi in dm)
                self.v[n] = np.zeros(shape, dtype=np.int64 if ty == "i" else np.float64, order="F")


# ------------------------------------------------------------------------------------------------
class Unit:
    def __init__(self, name, args, stmts, path):
        self.name, self.args, self.stmts, self.path = name, args, stmts, path


def split_units(path):
    units, cur = {}, None
    for ln, s in read_statements(path):
        low = s.lower()
        m = re.match(r"subroutine\s+([a-z_][a-z0-9_]*)\s*(?:\((.*)\))?$", low)
        if m:
            cur = Unit(m.group(1), [a.strip() for a in m.group(2).split(",")] if m.group(2) else [], [], path)
            units[cur.name] = cur
            continue
        if cur is None:
            continue   # `program` etc.
        if re.match(r"end(\s+subroutine(\s+\w+)?)?$", low):
            cur = None
            continue
        cur.stmts.append((ln, s))
    return units


# one rank: a sum / maximum / broadcast over the communicator leaves its argument as it is (parallel_mpi.f:125-151)
NOOP_CALLS = {"exchange2d_mpi", "exchange3d_mpi", "order2d_mpi", "order3d_mpi", "psum0d_mpi", "sum0d_mpi", "max0d_mpi",
              "bcast0d_mpi", "finalize_mpi", "msg_print"}


class Translator:
    def __init__(self, G, units):
        self.G, self.units = G, units
        self.src = {}

    # -- expressions --------------------------------------------------------------------------
    def ref(self, name):
        if name in self.locals:
            return "l_" + name
        if name in self.G.v:
            if isinstance(self.G.v[name], np.ndarray):
                self.used_garrays.add(name)
                return "g_" + name
            if name in self.G.params or name in ("im_global", "jm_global", "kb", "im_local", "jm_local", "n_proc"):
                return repr(self.G.v[name])
            return f"G[{name!r}]"
        raise NameError(f"{self.unit.name}: unknown name {name!r}")

    def is_array(self, name):
        if name in self.locals:
            return self.locals[name][1] is not None
        return name in self.G.v and isinstance(self.G.v[name], np.ndarray)

    def lower_bounds(self, name):
        dm = self.locals[name][1] if name in self.locals else self.G.dims[name]
        return [lo for lo, hi in dm]

    def ex(self, e):
        k = e[0]
        if k == "num":
            t = e[1]
            t = re.sub(r"_\w+$", "", t)
            if "d" in t:
                return f"f8({float(t.replace('d', 'e'))!r})"
            if "." in t or "e" in t:
                return f"f4({float(t)!r})"
            return str(int(t))
        if k == "bool":
            return "True" if e[1] else "False"
        if k == "str":
            return repr(e[1])
        if k == "par":
            return "(" + self.ex(e[1]) + ")"
        if k == "id":
            return self.ref(e[1])
        if k == "neg":
            return "(-" + self.ex(e[1]) + ")"
        if k == "not":
            return "(not " + self.ex(e[1]) + ")"
        if k in ("and", "or"):
            return f"({self.ex(e[1])} {k} {self.ex(e[2])})"
        if k == "rel":
            return f"({self.ex(e[2])} {e[1]} {self.ex(e[3])})"
        if k == "bin":
            a, b = self.ex(e[2]), self.ex(e[3])
            if e[1] == "/":
                return f"_div({a}, {b})"
            return f"({a} {e[1]} {b})"
        if k == "pow":
            return f"_pow({self.ex(e[1])}, {self.ex(e[2])})"
        if k == "call":
            name, args = e[1], e[2]
            if self.is_array(name):
                return self.ref(name) + "[" + self.subscripts(name, args) + "]"
            if name in INTRINSICS:
                return INTRINSICS[name] + "(" + ", ".join(self.ex(a) for a in args) + ")"
            raise NameError(f"{self.unit.name}: unknown function or array {name!r}")
        raise SyntaxError(f"bad node {e!r}")

    def subscripts(self, name, args):
        lbs = self.lower_bounds(name)
        out = []
        for a, lb in zip(args, lbs):
            off = self.ex(parse_expr(lb)) if lb else "1"
            if a[0] == "sec":
                lo = f"({self.ex(a[1])})-({off})" if a[1] is not None else ""
                hi = f"({self.ex(a[2])})-({off})+1" if a[2] is not None else ""
                out.append(f"{lo}:{hi}")
            else:
                out.append(f"({self.ex(a)})-{off}" if off != "1" else self.sub1(a))
        return ", ".join(out)

    def sub1(self, a):   # index - 1, folded for the common i, i+1, i-1 forms
        if a[0] == "id":
            return f"{self.ref(a[1])}-1"
        if a[0] == "num" and a[1].isdigit():
            return str(int(a[1]) - 1)
        if a[0] == "bin" and a[1] in "+-" and a[3][0] == "num" and a[3][1].isdigit() and a[2][0] == "id":
            d = int(a[3][1]) * (1 if a[1] == "+" else -1) - 1
            return f"{self.ref(a[2][1])}{d:+d}" if d else self.ref(a[2][1])
        return f"({self.ex(a)})-1"

    # -- statements ---------------------------------------------------------------------------
    def lhs_assign(self, lhs, rhs_code):
        if lhs[0] == "id":
            n = lhs[1]
            if self.is_array(n):
                return f"{self.ref(n)}[...] = {rhs_code}"
            ty = self.locals[n][0] if n in self.locals else self.G.types.get(n, "f8")
            conv = {"f8": "f8", "i": "int", "f4": "f4", "b": "bool", "f16": "f16", "c": "str"}[ty]
            tgt = self.ref(n)
            return f"{tgt} = {conv}({rhs_code})"
        if lhs[0] == "call":
            n = lhs[1]
            return f"{self.ref(n)}[{self.subscripts(n, lhs[2])}] = {rhs_code}"
        raise SyntaxError(f"bad assignment target {lhs!r}")

    def simple(self, s):
        """one non-block statement -> python line(s)"""
        low = s.lower()
        if low == "return":
            return ["return __RET__"]
        if low.startswith(("write", "print", "format", "open", "close", "read")):
            return ["pass"]
        if low.startswith("stop"):
            return ["raise RuntimeError('stop in %s')" % self.unit.name]
        m = re.match(r"call\s+([a-z_][a-z0-9_]*)\s*(?:\((.*)\))?$", low)
        if m:
            name = m.group(1)
            if name in NOOP_CALLS:
                return ["pass"]
            args = []
            if m.group(2):
                p = P(tokenize(m.group(2)))
                while True:
                    args.append(p.arg())
                    if p.peek()[1] == ",":
                        p.next(); continue
                    break
            self.called.add(name)
            return [f"R[{name!r}](" + ", ".join(self.ex(a) for a in args) + ")"]
        # assignment: split at the top-level '='
        depth = 0
        for p, ch in enumerate(low):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and low[p + 1:p + 2] != "=" and low[p - 1:p] not in "<>/=":
                lhs = parse_expr(low[:p])
                rhs = parse_expr(low[p + 1:])
                return [self.lhs_assign(lhs, self.ex(rhs))]
        raise SyntaxError(f"{self.unit.path}: cannot translate {s!r}")

    def translate(self, unit):
        self.unit = unit
        self.locals, self.used_garrays, self.called = {}, set(), set()
        body, decl_lines, data_lines = [], [], []
        ind = 1
        args = list(unit.args)
        for ln, s in unit.stmts:
            low = s.lower()
            pad = "    " * ind
            try:
                if low.startswith(("implicit", "include", "save", "external", "intrinsic")):
                    continue
                d = parse_decl(s) if not re.match(r"[a-z_0-9]+\s*(\(.*\))?\s*=", low) else None
                if d:
                    ty, names = d
                    for n, dm in names:
                        self.locals[n] = (ty, dm)
                    continue
                if low.startswith("parameter"):
                    inner = low[low.index("(") + 1:low.rindex(")")]
                    for item in split_top(inner):
                        k, v = item.split("=")
                        data_lines.append(self.lhs_assign(("id", k.strip()), self.ex(parse_expr(v))))
                    continue
                if low.startswith("data"):
                    for grp in re.findall(r"([a-z_0-9,\s]+)/([^/]*)/", low[4:]):
                        names = [x.strip() for x in grp[0].strip(" ,").split(",")]
                        vals = split_top(grp[1])
                        if len(names) == 1 and len(vals) > 1:      # array initialiser
                            data_lines.append(f"{self.ref(names[0])}[...] = [" + ", ".join(self.ex(parse_expr(v)) for v in vals) + "]")
                        else:
                            for n, v in zip(names, vals):
                                data_lines.append(self.lhs_assign(("id", n), self.ex(parse_expr(v))))
                    continue
                m = re.match(r"do\s+([a-z_][a-z0-9_]*)\s*=\s*(.*)$", low)
                if m:
                    parts = split_top(m.group(2))
                    a, b = self.ex(parse_expr(parts[0])), self.ex(parse_expr(parts[1]))
                    v = self.ref(m.group(1))
                    if len(parts) == 3:
                        st = self.ex(parse_expr(parts[2]))
                        body.append(f"{pad}for {v} in range({a}, ({b}) + (1 if ({st}) > 0 else -1), {st}):")
                    else:
                        body.append(f"{pad}for {v} in range({a}, ({b}) + 1):")
                    ind += 1
                    continue
                if re.match(r"end\s*do$", low) or re.match(r"end\s*if$", low):
                    ind -= 1
                    continue
                m = re.match(r"(else\s*)?if\s*\((.*)\)\s*then$", low)
                if m:
                    cond = self.ex(parse_expr(m.group(2)))
                    if m.group(1):
                        body.append(f"{'    ' * (ind - 1)}elif {cond}:")
                    else:
                        body.append(f"{pad}if {cond}:")
                        ind += 1
                    body.append(f"{'    ' * ind}pass")
                    continue
                if low == "else":
                    body.append(f"{'    ' * (ind - 1)}else:")
                    body.append(f"{pad}pass")
                    continue
                if low.startswith("if") and re.match(r"if\s*\(", low):
                    # logical IF: find the matching parenthesis
                    p0 = low.index("(")
                    depth = 0
                    for p in range(p0, len(low)):
                        if low[p] == "(":
                            depth += 1
                        elif low[p] == ")":
                            depth -= 1
                            if depth == 0:
                                break
                    cond = self.ex(parse_expr(low[p0 + 1:p]))
                    body.append(f"{pad}if {cond}:")
                    for line in self.simple(s[p + 1:].strip()):
                        body.append(f"{pad}    {line}")
                    continue
                for line in self.simple(s):
                    body.append(pad + line)
            except Exception as ex:
                raise type(ex)(f"{unit.path}:{ln}: {s!r}: {ex}") from ex
        # prologue: dummy arguments, locals, global arrays
        pro = []
        for n, (ty, dm) in self.locals.items():
            if n in args:
                continue
            if dm is None:
                pro.append(f"    l_{n} = " + {"f8": "f8(0.)", "i": "0", "f4": "f4(0.)", "b": "False", "f16": "f16(0.)", "c": "''"}[ty])
            else:
                shape = ", ".join(f"({self.ex(parse_expr(hi))})-({self.ex(parse_expr(lo)) if lo else 1})+1" for lo, hi in dm)
                pro.append(f"    l_{n} = np.zeros(({shape},), dtype=np.float64 if {ty != 'i'!r} else np.int64, order='F')")
        for n in args:
            if n not in self.locals:
                raise NameError(f"{unit.name}: undeclared dummy argument {n}")
        garr = [f"    g_{n} = G[{n!r}]" for n in sorted(self.used_garrays)]
        head = f"def {unit.name}(" + ", ".join("l_" + a for a in args) + "):"
        # scalar dummy arguments are passed by value: their final values are handed back as a tuple
        ret = "(" + "".join(f"l_{a}, " for a in args if self.locals[a][1] is None) + ")"
        src = "\n".join([head] + garr + pro + ["    " + x for x in data_lines] + body + ["    return __RET__"])
        return src.replace("__RET__", ret)


CACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # built artefacts: git-ignored, travel with gpurun


class Reference:
    """The translated hot path bound to one COMMON state.  Where the reference source is present it is read and
    translated now; elsewhere (the GPU box) a translation built earlier for the same grid by
    `build_cache` (oracle/_ref/f77_<im>x<jm>x<kb>.pkl, a build artefact like a compiled .so) is loaded."""
    FILES = ("pom/solver.f", "pom/advance.f", "pom/bounds_forcing.f")

    @staticmethod
    def cache_path(im, jm, kb):
        return os.path.join(CACHE_DIR, f"f77_{im}x{jm}x{kb}.pkl")

    @staticmethod
    def available(im, jm, kb, root=REF_ROOT):
        return os.path.exists(os.path.join(root, "pom", "solver.f")) or os.path.exists(Reference.cache_path(im, jm, kb))

    def __init__(self, im, jm, kb, root=REF_ROOT):
        self.root = root
        self.R = {}
        self.src = {}
        self.externals = {}
        if not os.path.exists(os.path.join(root, "pom", "solver.f")):
            import pickle
            with open(self.cache_path(im, jm, kb), "rb") as fh:
                c = pickle.load(fh)
            self.G = c["G"]
            self.src = c["src"]
            self.units = dict.fromkeys(self.src)
            self.tr = None
            return
        self.G = Globals(os.path.join(root, "pom.h_dist"),
                         dict(im_global=im, jm_global=jm, kb=kb, im_local=im, jm_local=jm, n_proc=1))
        v = self.G.v
        v.update(im=im, imm1=im - 1, imm2=im - 2, jm=jm, jmm1=jm - 1, jmm2=jm - 2, kbm1=kb - 1, kbm2=kb - 2)
        v.update(my_task=0, master_task=0, n_west=-1, n_east=-1, n_south=-1, n_north=-1)
        self.units = {}
        for f in self.FILES:
            self.units.update(split_units(os.path.join(root, f)))
        self.tr = Translator(self.G, self.units)

    def routine(self, name):
        if name not in self.R:
            if name not in self.units:
                raise KeyError(f"the reference has no subroutine {name!r} in {self.FILES}")
            if name not in self.src:
                self.src[name] = self.tr.translate(self.units[name])
            env = dict(RUNTIME)
            env["G"] = self.G.v
            env["R"] = _Lazy(self)
            exec(compile(self.src[name], f"<reference {name}>", "exec"), env)
            self.R[name] = env[name]
        return self.R[name]

    def call(self, name, *args):
        with np.errstate(all="ignore"):
            return self.routine(name)(*args)

    def build_cache(self):
        """Translate every subroutine of the three files and store the result for this grid under oracle/_ref/."""
        import pickle
        for n in self.units:
            if n not in self.src:
                self.src[n] = self.tr.translate(self.units[n])
        os.makedirs(CACHE_DIR, exist_ok=True)
        with open(self.cache_path(self.G.v["im_local"], self.G.v["jm_local"], self.G.v["kb"]), "wb") as fh:
            pickle.dump({"G": self.G, "src": self.src}, fh)


class _Lazy(dict):
    def __init__(self, ref):
        self.ref = ref

    def __missing__(self, name):
        if name in self.ref.externals:      # harness stand-ins for the file readers (PnetCDF) the path calls
            return self.ref.externals[name]
        return self.ref.routine(name)


# ------------------------------------------------------------------------------------------------
class F77Ref:
    """One sub-domain of the reference model, executed from the reference's own source.  Same Python surface
    as oracle.pomo.Oracle / extpom_b200.PomGpu (load / get / put / set / step and the subroutine names), so the
    parity tests drive all three the same way."""

    def __init__(self, im, jm, kb, root=REF_ROOT):
        self.im, self.jm, self.kb = im, jm, kb
        self.ref = Reference(im, jm, kb, root)
        self.v = self.ref.G.v
        self.types = self.ref.G.types
        # restore_interior reads its target fields from a netCDF file (bounds_forcing.f:1039-1081); the harness
        # plays the file: every record holds (restore_t, restore_s), default the climatology
        self.restore_t = self.restore_s = None
        self.ref.externals["read_restore_ts_interior_pnetcdf"] = self._read_restore

    def _read_restore(self, n, kb, tr, sr):
        tr[...] = self.v["tclim"] if self.restore_t is None else self.restore_t
        sr[...] = self.v["sclim"] if self.restore_s is None else self.restore_s

    def close(self):
        pass

    # -- state I/O -------------------------------------------------------
    def set(self, name, val):
        if name not in self.v or isinstance(self.v[name], np.ndarray):
            raise KeyError(name)
        ty = self.types.get(name, "f8")
        self.v[name] = int(val) if ty == "i" else (bool(val) if ty == "b" else f8(val))

    def getc(self, name):
        return float(self.v[name])

    def load(self, state):
        for k, val in state["consts"].items():
            if k in self.v and not isinstance(self.v[k], np.ndarray):
                self.set(k, val)
        for k, a in state["fields"].items():
            if k in self.v and isinstance(self.v[k], np.ndarray):
                self.v[k][...] = a

    def get(self, name):
        return np.array(self.v[name], order="F", copy=True)

    def put(self, name, arr):
        self.v[name][...] = arr

    def _a(self, x):
        return self.v[x] if isinstance(x, str) else x

    # -- the reference's subroutine surface --------------------------------
    def step(self, iint, time=None):
        """advance.f:21-32 for internal step `iint`, with get_time's clock (advance.f:62-75)."""
        self.v["iint"] = int(iint)
        self.v["time"] = f8(self.v["dti"] * float(iint) / 86400.0 + self.v["time0"] if time is None else time)
        c = self.ref.call
        c("lateral_viscosity")
        c("mode_interaction")
        for iext in range(1, self.v["isplit"] + 1):
            self.v["iext"] = iext
            c("mode_external")
        c("mode_internal")

    def check_velocity(self):
        vaf = self.v["vaf"]
        return float(np.abs(vaf).max())

    def lateral_viscosity(self): self.ref.call("lateral_viscosity")
    def mode_interaction(self): self.ref.call("mode_interaction")

    def mode_external(self, iext):
        self.v["iext"] = int(iext); self.ref.call("mode_external")

    def mode_internal(self, iint):
        self.v["iint"] = int(iint); self.ref.call("mode_internal")

    def advave(self): self.ref.call("advave")
    def advct(self): self.ref.call("advct")
    def advu(self): self.ref.call("advu")
    def advv(self): self.ref.call("advv")
    def baropg(self): self.ref.call("baropg")
    def baropg_mcc(self): self.ref.call("baropg_mcc")
    def profq(self): self.ref.call("profq")
    def profu(self): self.ref.call("profu")
    def profv(self): self.ref.call("profv")
    def vertvl(self): self.ref.call("vertvl")
    def realvertvl(self): self.ref.call("realvertvl")
    def bcond(self, idx): self.ref.call("bcond", int(idx))
    def bcondorl(self, idx): self.ref.call("bcondorl", int(idx))
    def advq(self, qb, q, qf): self.ref.call("advq", self._a(qb), self._a(q), self._a(qf))
    def advt1(self, fb, f, fclim, ff): self.ref.call("advt1", self._a(fb), self._a(f), self._a(fclim), self._a(ff))
    def advt2(self, fb, f, fclim, ff): self.ref.call("advt2", self._a(fb), self._a(f), self._a(fclim), self._a(ff))
    def dens(self, si, ti, rhoo): self.ref.call("dens", self._a(si), self._a(ti), self._a(rhoo))
    def proft(self, f, wfsurf, fsurf, nbc): self.ref.call("proft", self._a(f), self._a(wfsurf), self._a(fsurf), int(nbc))
    def smol_adif(self, xm, ym, zw, ff): self.ref.call("smol_adif", self._a(xm), self._a(ym), self._a(zw), self._a(ff))

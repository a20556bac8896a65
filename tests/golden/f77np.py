"""f77np -- a small Fortran-77/90 subset interpreter on numpy.

TEST INFRASTRUCTURE (golden-vector generator), never imported by phasta_b200/.

Purpose: execute the UNMODIFIED reference sources under /root/reference
(phSolver/compressible/*.f, phSolver/common/*.f, common.h) in this container,
which has no Fortran compiler, so that golden input/output vectors for the
hot path come from the reference's own statements and not from a hand
restatement.  tests/golden/make_golden_f77.py drives it and commits the
vectors as .npz fixtures; the oracle (oracle/*.c) is then pinned against them.

Semantics implemented (what the hot-path sources use):
  * fixed-form source, `c`/`!` comments, continuation lines, tabs, labels;
  * `include "common.h"`: COMMON blocks, PARAMETERs and type statements are
    parsed from the reference's own common.h; every unit that includes it sees
    the same global storage (a COMMON variable assigned in one routine is seen
    by the next, exactly as in the reference);
  * implicit typing REAL*8 (a-h,o-z) / INTEGER (i-n), integer division,
    assignment conversion;
  * whole-array / array-section / vector-subscript expressions, WHERE /
    ELSEWHERE, DO (with labels), IF / ELSE IF, CYCLE / EXIT / RETURN / GOTO to
    a label in an enclosing block, ALLOCATE, derived-type `%` component refs;
  * CALL with pass-by-reference: array sections are numpy views, scalars are
    copied back, an array element passed to an array dummy is sequence-
    associated (flat column-major view from that element), dummy arrays are
    re-shaped views when the declared shape differs from the actual one;
  * automatic (local) REAL arrays are filled with NaN so that a read of an
    uninitialised local shows up in the output instead of passing silently.
Arithmetic is IEEE double evaluated strictly left-to-right with the Fortran
precedence rules (what an unoptimised build does).
"""
import keyword
import math
import os
import re

import numpy as np

# ----------------------------------------------------------------------------
# source reader
# ----------------------------------------------------------------------------
_CONT = re.compile(r"^(     [^ 0\t]|\t[1-9]| {0,4}&)")


def _strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == "!":
            break
        out.append(ch)
    return "".join(out).rstrip()


def read_statements(path):
    """-> list of (label or None, statement text, line number)."""
    stmts = []
    with open(path, errors="replace") as f:
        raw = f.read().split("\n")
    cur, cur_no = None, 0
    for no, line in enumerate(raw, 1):
        if not line.strip():
            continue
        if line[0] in "cC*!" or line.lstrip().startswith("!"):
            continue
        if line.lstrip().startswith("#"):
            continue
        line = _strip_comment(line)
        if not line.strip():
            continue
        m = _CONT.match(line)
        if m and cur is not None:
            body = line[m.end():] if not line.lstrip().startswith("&") else line.lstrip()[1:]
            cur += " " + body.strip()
            continue
        if cur is not None:
            stmts.append((cur, cur_no))
        cur, cur_no = line.strip(), no
        # free-form trailing '&' continuation is not used by these sources
    if cur is not None:
        stmts.append((cur, cur_no))
    out = []
    for s, no in stmts:
        m = re.match(r"^(\d+)\s+(.*)$", s)
        label = None
        if m:
            label, s = int(m.group(1)), m.group(2)
        for part in _split_semicolon(s):
            out.append((label, part, no))
            label = None
    return out


def _split_semicolon(s):
    if ";" not in s:
        return [s]
    parts, cur, q = [], [], None
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == ";":
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return [p for p in parts if p]


# ----------------------------------------------------------------------------
# tokenizer / expression compiler
# ----------------------------------------------------------------------------
_DOTOPS = "eq|ne|lt|le|gt|ge|and|or|not|eqv|neqv|true|false"
_TOK = re.compile(
    r"\s*(?:"
    r"(?P<int_dot>\d+(?=\.(?:%s)\.))" % _DOTOPS +
    r"|(?P<num>(?:\d+\.?\d*|\.\d+)(?:[dDeE][+-]?\d+)?(?:_\d+)?)"
    r"|(?P<dot>\.(?:%s)\.)" % _DOTOPS +
    r"|(?P<id>[A-Za-z_][A-Za-z0-9_]*)"
    r"|(?P<str>'(?:[^']|'')*'|\"(?:[^\"]|\"\")*\")"
    r"|(?P<op>\*\*|//|==|/=|<=|>=|=>|::|[-+*/(),:=<>%])"
    r")", re.I)


_PYKW = set(keyword.kwlist) - {"if", "else", "return", "continue", "while"}
_DOTFIX = re.compile(r"\.\s*(%s)\s*\.(?!\d)" % _DOTOPS, re.I)


def tokenize(s):
    toks, pos = [], 0
    s = _DOTFIX.sub(lambda m: "." + m.group(1) + ".", s.rstrip())
    while pos < len(s):
        m = _TOK.match(s, pos)
        if not m or m.end() == pos:
            raise SyntaxError("cannot tokenize %r at %d" % (s, pos))
        pos = m.end()
        if m.group("int_dot"):
            toks.append(("num", m.group("int_dot")))
        elif m.group("num"):
            toks.append(("num", m.group("num")))
        elif m.group("dot"):
            toks.append(("op", m.group("dot").lower()))
        elif m.group("id"):
            name = m.group("id").lower()
            toks.append(("id", name + "_" if name in _PYKW else name))
        elif m.group("str"):
            t = m.group("str")
            q = t[0]
            toks.append(("str", t[1:-1].replace(q + q, q)))
        else:
            toks.append(("op", m.group("op")))
    return toks


class ExprParser:
    """Fortran expression -> python source using the runtime helpers."""

    def __init__(self, toks, pos=0):
        self.t, self.p = toks, pos

    def peek(self):
        return self.t[self.p] if self.p < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.p += 1
        return tok

    def accept(self, val):
        if self.peek() == ("op", val):
            self.p += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise SyntaxError("expected %r got %r in %r" % (val, self.peek(), self.t))

    def expr(self):
        return self.eqv()

    def eqv(self):
        a = self.or_()
        while self.peek() in (("op", ".eqv."), ("op", ".neqv.")):
            op = self.next()[1]
            b = self.or_()
            a = "_eqv(%s,%s)" % (a, b) if op == ".eqv." else "_neqv(%s,%s)" % (a, b)
        return a

    def or_(self):
        a = self.and_()
        while self.accept(".or."):
            a = "_or(%s,%s)" % (a, self.and_())
        return a

    def and_(self):
        a = self.not_()
        while self.accept(".and."):
            a = "_and(%s,%s)" % (a, self.not_())
        return a

    def not_(self):
        if self.accept(".not."):
            return "_not(%s)" % self.not_()
        return self.rel()

    _REL = {".eq.": "_eq", "==": "_eq", ".ne.": "_ne", "/=": "_ne", ".lt.": "<", "<": "<",
            ".le.": "<=", "<=": "<=", ".gt.": ">", ">": ">", ".ge.": ">=", ">=": ">="}

    def rel(self):
        a = self.concat()
        tok = self.peek()
        if tok[0] == "op" and tok[1] in self._REL:
            self.next()
            b = self.concat()
            f = self._REL[tok[1]]
            return "%s(%s,%s)" % (f, a, b) if f.startswith("_") else "(%s%s%s)" % (a, f, b)
        return a

    def concat(self):
        a = self.add()
        while self.accept("//"):
            a = "(%s+%s)" % (a, self.add())
        return a

    def add(self):
        if self.accept("-"):
            a = "(-%s)" % self.mul()
        elif self.accept("+"):
            a = self.mul()
        else:
            a = self.mul()
        while True:
            if self.accept("+"):
                a = "(%s+%s)" % (a, self.mul())
            elif self.accept("-"):
                a = "(%s-%s)" % (a, self.mul())
            else:
                return a

    def mul(self):
        a = self.pow()
        while True:
            if self.accept("*"):
                a = "(%s*%s)" % (a, self.pow())
            elif self.peek() == ("op", "/") and self.p + 1 < len(self.t) and self.t[self.p + 1] == ("op", ")"):
                return a  # closing of an array constructor (/ ... /)
            elif self.accept("/"):
                a = "_div(%s,%s)" % (a, self.pow())
            else:
                return a

    def pow(self):
        a = self.primary()
        if self.accept("**"):
            # right associative; unary minus allowed in the exponent
            if self.accept("-"):
                b = "(-%s)" % self.pow()
            else:
                b = self.pow()
            return "_pow(%s,%s)" % (a, b)
        return a

    def arglist(self):
        """after '(' : list of python sources; sections become _sl(...)"""
        args = []
        if self.accept(")"):
            return args
        while True:
            args.append(self.arg())
            if self.accept(","):
                continue
            self.expect(")")
            return args

    def arg(self):
        # keyword argument (dim=1) -> positional
        if self.peek()[0] == "id" and self.p + 1 < len(self.t) and self.t[self.p + 1] == ("op", "=") \
                and (self.p + 2 >= len(self.t) or self.t[self.p + 2] != ("op", "=")):
            self.p += 2
        lo = hi = st = None
        if self.peek() != ("op", ":"):
            lo = self.expr()
            if self.peek() != ("op", ":"):
                return lo
        self.expect(":")
        if self.peek() not in (("op", ","), ("op", ")"), ("op", ":")):
            hi = self.expr()
        if self.accept(":"):
            st = self.expr()
        return "_sl(%s,%s,%s)" % (lo, hi, st)

    def primary(self):
        kind, val = self.next()
        if kind == "num":
            v = re.sub(r"_\d+$", "", val)
            if re.search(r"[.dDeE]", v):
                return repr(float(re.sub("[dD]", "e", v)))
            return str(int(v))
        if kind == "str":
            return repr(val)
        if kind == "op" and val == "(" and self.peek() == ("op", "/"):
            self.next()
            items = []
            while True:
                items.append(self.expr())
                if self.accept(","):
                    continue
                self.expect("/")
                self.expect(")")
                return "np.array([%s])" % ",".join(items)
        if kind == "op" and val == "(":
            a = self.expr()
            if self.accept(","):  # complex literal -- not used
                raise SyntaxError("complex literal")
            self.expect(")")
            return "(%s)" % a
        if kind == "op" and val == ".true.":
            return "True"
        if kind == "op" and val == ".false.":
            return "False"
        if kind == "op" and val == "-":
            return "(-%s)" % self.primary()
        if kind == "id" and val.endswith("_") and self.peek() is not None and self.peek()[0] == "str":
            return repr(self.next()[1])          # kind-prefixed character literal: c_char_'integer'
        if kind == "id" and val == "c_loc" and self.peek() == ("op", "(") and self.p + 2 < len(self.t) \
                and self.t[self.p + 1][0] == "id" and self.t[self.p + 2] == ("op", ")"):
            name = self.t[self.p + 1][1]             # c_loc(name): the array itself, or a reference to the scalar
            self.p += 3
            return "_cloc(_E,%r)" % name
        if kind == "id":
            src = val
            if self.accept("("):
                args = self.arglist()
                if val in FUNCTION_UNITS:
                    # actual arguments of a function subprogram: an array element is
                    # sequence-associated (flat view from that element)
                    args = [re.sub(r"^_ref\(", "_elemview(", a) if re.match(r"^_ref\(_E,'\w+',\([^:]*\)\)$", a)
                            and "_sl(" not in a else a for a in args]
                src = "_ref(_E,%r,(%s))" % (val, "".join(a + "," for a in args))
            while self.accept("%"):
                k, f = self.next()
                src = "%s.%s" % (src, f)
                if self.accept("("):
                    args = self.arglist()
                    src = "_idx(%s,(%s))" % (src, "".join(a + "," for a in args))
            return src
        raise SyntaxError("unexpected token %r in %r" % ((kind, val), self.t))


_code_cache = {}


def compile_expr(src):
    c = _code_cache.get(src)
    if c is None:
        c = compile(src, "<f77>", "eval")
        _code_cache[src] = c
    return c


def parse_expr_tokens(toks):
    p = ExprParser(toks)
    src = p.expr()
    if p.p != len(toks):
        raise SyntaxError("trailing tokens %r in %r" % (toks[p.p:], toks))
    return src


# ----------------------------------------------------------------------------
# runtime helpers (globals of every eval)
# ----------------------------------------------------------------------------
_INTT = (int, np.integer)


def _isint(a):
    if isinstance(a, (bool, np.bool_)):
        return False
    if isinstance(a, _INTT):
        return True
    return isinstance(a, np.ndarray) and a.dtype.kind in "iu"


def _div(a, b):
    if _isint(a) and _isint(b):
        q = np.abs(a) // np.abs(b)
        q = q * np.sign(a) * np.sign(b)
        return int(q) if np.ndim(q) == 0 else q.astype(np.int64)
    return a / b


def _pow(a, b):
    if _isint(a) and _isint(b) and np.ndim(a) == 0 and np.ndim(b) == 0:
        return int(a) ** int(b) if b >= 0 else 0
    return a ** b


def _strip(a):
    return a.rstrip() if isinstance(a, str) else a


def _eq(a, b):
    return _strip(a) == _strip(b)


def _ne(a, b):
    return _strip(a) != _strip(b)


def _and(a, b):
    if np.ndim(a) == 0 and np.ndim(b) == 0:
        return bool(a) and bool(b)
    return np.logical_and(a, b)


def _or(a, b):
    if np.ndim(a) == 0 and np.ndim(b) == 0:
        return bool(a) or bool(b)
    return np.logical_or(a, b)


def _not(a):
    return (not bool(a)) if np.ndim(a) == 0 else np.logical_not(a)


def _eqv(a, b):
    return _not(_neqv(a, b))


def _neqv(a, b):
    return np.logical_xor(a, b) if (np.ndim(a) or np.ndim(b)) else (bool(a) != bool(b))


class _Sl:
    __slots__ = ("lo", "hi", "st")

    def __init__(self, lo, hi, st):
        self.lo, self.hi, self.st = lo, hi, st


def _sl(lo, hi, st):
    return _Sl(lo, hi, st)


def _conv_index(args, lbounds):
    out = []
    for k, a in enumerate(args):
        lb = lbounds[k] if lbounds and k < len(lbounds) else 1
        if isinstance(a, _Sl):
            lo = None if a.lo is None else int(a.lo) - lb
            hi = None if a.hi is None else int(a.hi) - lb + 1
            st = None if a.st is None else int(a.st)
            if st is not None and st < 0:
                hi = None if a.hi is None else int(a.hi) - lb - 1
                if hi is not None and hi < 0:
                    hi = None
            out.append(slice(lo, hi, st))
        elif isinstance(a, np.ndarray):
            out.append(a.astype(np.int64) - lb)
        else:
            i = int(a) - lb
            if i < 0:
                raise IndexError("Fortran index %r below lower bound %d" % (a, lb))
            out.append(i)
    return tuple(out)


def _idx(obj, args):
    if isinstance(obj, (list, tuple)):
        return obj[int(args[0]) - 1]
    return obj[_conv_index(args, None)]


def _sign(a, b):
    return np.where(np.asarray(b) >= 0, np.abs(a), -np.abs(a)) if (np.ndim(a) or np.ndim(b)) \
        else (abs(a) if b >= 0 else -abs(a))


def _minmax(fn):
    def f(*a):
        r = a[0]
        for x in a[1:]:
            r = fn(r, x)
        return r
    return f


def _btest(i, pos):
    r = (np.asarray(i) >> pos) & 1
    return bool(r) if np.ndim(r) == 0 else r.astype(bool)


def _ibits(i, pos, ln):
    r = (np.asarray(i) >> pos) & ((1 << ln) - 1)
    return int(r) if np.ndim(r) == 0 else r


def _int(a):
    r = np.trunc(a)
    return int(r) if np.ndim(r) == 0 else r.astype(np.int64)


def _nint(a):
    r = np.where(np.asarray(a) >= 0, np.floor(np.asarray(a) + 0.5), -np.floor(-np.asarray(a) + 0.5))
    return int(r) if np.ndim(r) == 0 else r.astype(np.int64)


def _mod(a, b):
    r = np.fmod(a, b)
    if _isint(a) and _isint(b):
        return int(r) if np.ndim(r) == 0 else r.astype(np.int64)
    return r


def _sum(a, dim=None):
    if dim is None:
        # Fortran SUM of a whole array: sequential in array-element order
        flat = np.asarray(a).reshape(-1, order="F")
        if flat.size == 0:
            return 0.0 if flat.dtype.kind == "f" else 0
        s = np.cumsum(flat)[-1]          # cumsum accumulates sequentially
        return float(s) if flat.dtype.kind == "f" else int(s)
    return np.add.reduce(a, axis=int(dim) - 1)


def _float(a):
    return float(a) if np.ndim(a) == 0 else np.asarray(a, dtype=np.float64)


INTRINSICS = {
    "sqrt": np.sqrt, "dsqrt": np.sqrt, "abs": np.abs, "dabs": np.abs, "exp": np.exp, "dexp": np.exp,
    "log": np.log, "dlog": np.log, "log10": np.log10, "sin": np.sin, "cos": np.cos, "tan": np.tan,
    "atan": np.arctan, "atan2": np.arctan2, "acos": np.arccos, "asin": np.arcsin, "tanh": np.tanh,
    "max": _minmax(np.maximum), "min": _minmax(np.minimum), "dmax1": _minmax(np.maximum),
    "dmin1": _minmax(np.minimum), "amax1": _minmax(np.maximum), "amin1": _minmax(np.minimum),
    "sign": _sign, "dsign": _sign, "btest": _btest, "ibits": _ibits, "int": _int, "nint": _nint,
    "mod": _mod, "sum": _sum, "maxval": lambda a: np.max(a), "minval": lambda a: np.min(a),
    "dble": _float, "float": _float, "real": _float, "dfloat": _float,
    "size": lambda a, d=None: a.size if d is None else a.shape[int(d) - 1],
    "any": lambda a: bool(np.any(a)), "all": lambda a: bool(np.all(a)),
    "count": lambda a: int(np.count_nonzero(a)),
    "iand": lambda a, b: a & b, "ior": lambda a, b: a | b,
    "secs": lambda *a: 0.0, "tmr": lambda *a: 0.0, "tmrc": lambda *a: 0.0,
    "isnan": lambda a: np.isnan(a),
    "transpose": lambda a: np.asfortranarray(np.transpose(a)),
    "dot_product": lambda a, b: _sum(a * b),
    "merge": lambda a, b, m: np.where(m, a, b),
    "trim": lambda s: s.rstrip(), "len": len,
    "c_loc": lambda a: a, "char": lambda i: chr(int(i)),     # iso_c_binding: the array itself stands for its address
    "minloc": lambda a: np.array([int(np.argmin(a)) + 1]), "maxloc": lambda a: np.array([int(np.argmax(a)) + 1]),
}


def _ref(E, name, args):
    try:
        obj = E[name]
    except KeyError:
        obj = None
    if isinstance(obj, np.ndarray):
        return obj[_conv_index(args, E.lbounds(name))]
    if isinstance(obj, (list, tuple)):
        return obj[int(args[0]) - 1]
    if obj is None and name in E.prog.units:     # function subprogram: result = variable named like it
        return E.prog.call(name, *args)[name]
    if obj is None or callable(obj):
        f = obj if callable(obj) else INTRINSICS.get(name)
        if f is None:
            raise NameError("f77np: unknown array or function %r" % name)
        return f(*args)
    if isinstance(obj, str):  # substring
        a = args[0]
        return obj[int(a.lo) - 1:int(a.hi)]
    raise TypeError("f77np: %r is a scalar (%r) but is referenced with arguments" % (name, obj))


class ScalarRef:
    """c_loc(scalar): a reference a python stub can store through (phio_readheader writing a COMMON scalar)"""

    def __init__(self, E, name):
        self.E, self.name = E, name

    def set(self, v):
        self.E.prog._assign(self.E, self.name, None, v)

    def get(self):
        return self.E[self.name]


def _cloc(E, name):
    obj = E[name]
    return obj if isinstance(obj, np.ndarray) else ScalarRef(E, name)


def _elemview(E, name, args):
    try:
        obj = E[name]
    except KeyError:
        obj = None
    if isinstance(obj, np.ndarray) and all(not isinstance(a, (_Sl, np.ndarray)) for a in args):
        idx = _conv_index(args, E.lbounds(name))
        off = int(np.ravel_multi_index(idx, obj.shape, order="F"))
        return obj.reshape(-1, order="F")[off:]
    return _ref(E, name, args)


FUNCTION_UNITS = set()


def scan_functions(path):
    """register the function subprograms of a file before anything is parsed"""
    for label, text, no in read_statements(path):
        m = re.match(r"^(?:[\w*() ]+\s+)?function\s+(\w+)\s*\(", text, re.I)
        if m and not re.match(r"^\s*end", text, re.I):
            FUNCTION_UNITS.add(m.group(1).lower())


HELPERS = dict(_cloc=_cloc, _elemview=_elemview, _ref=_ref, _idx=_idx, _sl=_sl, _div=_div, _pow=_pow, _eq=_eq, _ne=_ne, _and=_and,
               _or=_or, _not=_not, _eqv=_eqv, _neqv=_neqv, np=np, math=math)


# ----------------------------------------------------------------------------
# program representation
# ----------------------------------------------------------------------------
class Decl:
    __slots__ = ("name", "dims", "typ", "alloc")

    def __init__(self, name, dims=None, typ=None, alloc=False):
        self.name, self.dims, self.typ, self.alloc = name, dims, typ, alloc


class Unit:
    def __init__(self, name, args, path):
        self.name, self.args, self.path = name, args, path
        self.decls = {}      # name -> Decl
        self.order = []      # declaration order
        self.body = []
        self.uses_common = False
        self.params = []     # (name, code)


class CompPtr:
    """one element of a module array of derived-type pointers (x(i)%p)"""
    p = None


class _Return(Exception):
    pass


class _Cycle(Exception):
    pass


class _Exit(Exception):
    pass


class _Goto(Exception):
    def __init__(self, label):
        self.label = label


_TYPES = ("real", "integer", "logical", "character", "double", "complex")
_IGNORED = ("write", "print", "read", "open", "close", "format", "implicit", "external", "save",
            "intrinsic", "data", "use", "rewind", "flush", "equivalence", "nullify", "interface", "type")


def _split_top(toks, sep=","):
    parts, cur, depth = [], [], 0
    for t in toks:
        if t == ("op", "("):
            depth += 1
        elif t == ("op", ")"):
            depth -= 1
        if depth == 0 and t == ("op", sep):
            parts.append(cur)
            cur = []
        else:
            cur.append(t)
    parts.append(cur)
    return parts


def _parse_entities(toks):
    """a(n,m), b, c(0:k)*8  -> [(name, dims or None)], dims = [(lo_code|None, hi_code|'*'|':')]"""
    ents = []
    for ent in _split_top(toks):
        if not ent:
            continue
        if ent[0][0] != "id":
            raise SyntaxError("bad entity %r" % (ent,))
        name = ent[0][1]
        dims = None
        if len(ent) > 1 and ent[1] == ("op", "("):
            depth, j = 0, 1
            for j in range(1, len(ent)):
                if ent[j] == ("op", "("):
                    depth += 1
                elif ent[j] == ("op", ")"):
                    depth -= 1
                    if depth == 0:
                        break
            dims = []
            for d in _split_top(ent[2:j]):
                lohi = _split_top(d, ":")
                if len(lohi) == 1:
                    if d == [("op", "*")]:
                        dims.append((None, "*"))
                    else:
                        dims.append((None, parse_expr_tokens(d)))
                else:
                    lo = parse_expr_tokens(lohi[0]) if lohi[0] else None
                    if not lohi[1]:
                        dims.append((lo, ":"))
                    elif lohi[1] == [("op", "*")]:
                        dims.append((lo, "*"))
                    else:
                        dims.append((lo, parse_expr_tokens(lohi[1])))
        ents.append((name, dims))
    return ents


class Program:
    def __init__(self, include_dirs, modules=None, stubs=None, nan_locals=True):
        self.include_dirs = include_dirs
        self.units = {}
        self.G = {}            # COMMON + PARAMETER storage
        self.gl_lbounds = {}
        self.gtypes = {}
        self.M = dict(modules or {})   # module variables supplied by the driver
        self.stubs = dict(stubs or {})  # python callables replacing subroutines
        self.nan_locals = nan_locals
        self.common_loaded = False
        self.trace = None

    # ---- loading ----------------------------------------------------------
    def load(self, path):
        stmts = read_statements(path)
        i = 0
        while i < len(stmts):
            label, text, no = stmts[i]
            toks = tokenize(text)
            head = toks[0][1] if toks else ""
            is_sub = head == "subroutine" or (head in _TYPES and any(t == ("id", "function") for t in toks[:6])) \
                or head == "function"
            if head == "module" and len(toks) == 2:
                # skip module blocks (declarations come from the driver)
                while not re.match(r"^\s*end\s*module", stmts[i][1], re.I):
                    i += 1
                i += 1
                continue
            if not is_sub:
                i += 1
                continue
            k = [t for t in toks].index(("id", "subroutine")) if ("id", "subroutine") in toks else \
                [t for t in toks].index(("id", "function"))
            name = toks[k + 1][1]
            args = [t[1] for t in toks[k + 2:] if t[0] == "id"]
            unit = Unit(name, args, path)
            i += 1
            flat = []
            while i < len(stmts):
                label, text, no = stmts[i]
                if re.match(r"^end(\s*(subroutine|function)(\s+\w+)?)?\s*$", text.strip(), re.I):
                    i += 1
                    break
                flat.append((label, text, no))
                i += 1
            self._build_unit(unit, flat)
            self.units[name] = unit

    def _load_common(self, fname):
        for d in self.include_dirs:
            p = os.path.join(d, fname)
            if os.path.exists(p):
                break
        else:
            return False
        if self.common_loaded:
            return True
        self.common_loaded = True
        pend_dims = {}
        common_names = []
        for label, text, no in read_statements(p):
            toks = tokenize(text)
            head = toks[0][1]
            if head == "parameter":
                for ent in _split_top(toks[2:-1]):
                    name = ent[0][1]
                    val = eval(compile_expr(parse_expr_tokens(ent[2:])), HELPERS, _Env(self, {}, None))
                    if not isinstance(val, str) and name[0] in "ijklmn" and self.gtypes.get(name) != "real":
                        val = int(val)
                    self.G[name] = val
            elif head == "common":
                j = 1
                if toks[1] == ("op", "/"):
                    j = 4
                for name, dims in _parse_entities(toks[j:]):
                    common_names.append(name)
                    if dims:
                        pend_dims[name] = dims
            elif head in _TYPES:
                typ, rest = self._split_type(toks)
                for name, dims in _parse_entities(rest):
                    self.gtypes[name] = typ
                    if dims:
                        pend_dims[name] = dims
            elif head == "dimension":
                for name, dims in _parse_entities(toks[1:]):
                    pend_dims[name] = dims
        env = _Env(self, {}, None)
        for name in common_names:
            typ = self.gtypes.get(name) or ("integer" if name[0] in "ijklmn" else "real")
            if name in pend_dims:
                shape, lbs = [], []
                for lo, hi in pend_dims[name]:
                    l = 1 if lo is None else int(eval(compile_expr(lo), HELPERS, env))
                    h = int(eval(compile_expr(hi), HELPERS, env))
                    shape.append(h - l + 1)
                    lbs.append(l)
                dt = np.int64 if typ == "integer" else (np.bool_ if typ == "logical" else np.float64)
                if typ == "character":
                    self.G[name] = [""] * shape[0]
                else:
                    self.G[name] = np.zeros(shape, dtype=dt, order="F")
                if any(l != 1 for l in lbs):
                    self.gl_lbounds[name] = tuple(lbs)
            else:
                self.G[name] = 0 if typ == "integer" else (False if typ == "logical" else ("" if typ == "character" else 0.0))
        return True

    @staticmethod
    def _split_type(toks):
        """type-spec tokens -> (typ, entity tokens)"""
        typ = toks[0][1]
        if typ == "double":
            typ = "real"
            j = 2
        else:
            j = 1
        if ("op", "::") in toks:
            k = toks.index(("op", "::"))
            attrs = toks[j:k]
            ent = toks[k + 1:]
            alloc = any(t == ("id", "allocatable") or t == ("id", "pointer") for t in attrs)
            if ("id", "dimension") in attrs:
                # F90 attribute form `double precision, dimension(npro,nsd) :: a, b`: give the shape to
                # every entity that does not carry its own
                a0 = attrs.index(("id", "dimension")) + 1
                depth, a1 = 0, a0
                for a1 in range(a0, len(attrs)):
                    if attrs[a1] == ("op", "("):
                        depth += 1
                    elif attrs[a1] == ("op", ")"):
                        depth -= 1
                        if depth == 0:
                            break
                shape = attrs[a0:a1 + 1]
                out = []
                for e in _split_top(ent):
                    if out:
                        out.append(("op", ","))
                    out += e if ("op", "(") in e else e + shape
                ent = out
            return (typ + ("+alloc" if alloc else "")), ent
        # real*8 x / character*8 code / character(8) x / integer*8
        if j < len(toks) and toks[j] == ("op", "*"):
            j += 2
        elif j < len(toks) and toks[j] == ("op", "(") and typ == "character":
            depth = 0
            while True:
                if toks[j] == ("op", "("):
                    depth += 1
                elif toks[j] == ("op", ")"):
                    depth -= 1
                    if depth == 0:
                        j += 1
                        break
                j += 1
        return typ, toks[j:]

    def _build_unit(self, unit, flat):
        """declarations + nested statement blocks"""
        pos = [0]

        def decl(name, dims, typ):
            alloc = False
            if typ and typ.endswith("+alloc"):
                typ, alloc = typ[:-6], True
            d = unit.decls.get(name)
            if d is None:
                d = Decl(name)
                unit.decls[name] = d
                unit.order.append(name)
            if dims is not None:
                d.dims = dims
            if typ is not None:
                d.typ = typ
            d.alloc = d.alloc or alloc

        def parse_block(terminators):
            block = []
            while pos[0] < len(flat):
                label, text, no = flat[pos[0]]
                toks = tokenize(text)
                if not toks:
                    pos[0] += 1
                    continue
                head = toks[0][1]
                h2 = toks[1][1] if len(toks) > 1 else ""
                key = head + h2 if head == "end" and h2 in ("do", "if", "where", "select") else head
                if head == "else" and h2 == "if":
                    key = "elseif"
                if key in terminators:
                    return block, key, toks, label
                # labelled DO terminator
                if label is not None and ("label", label) in terminators:
                    # the labelled statement itself belongs to the loop body
                    stmt = self._parse_stmt(unit, toks, text, no, parse_block, pos, decl, label)
                    if stmt is not None:
                        block.append(stmt)
                    return block, ("label", label), toks, label
                pos[0] += 1
                stmt = self._parse_stmt(unit, toks, text, no, parse_block, pos, decl, label)
                if stmt is not None:
                    block.append(stmt)
            return block, None, None, None

        unit.body, _, _, _ = parse_block(())

    def _parse_stmt(self, unit, toks, text, no, parse_block, pos, decl, label):
        head = toks[0][1]
        where = "%s:%d" % (os.path.basename(unit.path), no)
        lab = ("label", label, where) if label is not None else None

        def with_label(stmt):
            if lab is None:
                return stmt
            return ("seq", [lab, stmt] if stmt is not None else [lab], where)

        if toks[0][0] == "id" and len(toks) > 1 and self._is_assignment(toks):
            return with_label(self._parse_assign(toks, where))
        if head == "include":
            fname = toks[1][1]
            if self._load_common(fname):
                unit.uses_common = True
            return None
        if head == "write" and len(toks) == 6 and toks[1] == ("op", "(") and toks[2][0] == "id" \
                and toks[3] == ("op", ",") and toks[4][0] == "str" and toks[5] == ("op", ")"):
            # internal write of a literal-only format into a character variable: write (name,"('text')")
            m = re.match(r"^\(\s*'((?:[^']|'')*)'\s*\)$", toks[4][1])
            if m:
                lit = m.group(1).replace("''", "'")
                return with_label(self._parse_assign([toks[2], ("op", "="), ("str", lit)], where))
        if head in _IGNORED:
            if head == "use":
                return None
            return with_label(None)
        if head == "parameter":
            for ent in _split_top(toks[2:-1]):
                unit.params.append((ent[0][1], parse_expr_tokens(ent[2:])))
            return None
        if head == "dimension":
            for name, dims in _parse_entities(toks[1:]):
                decl(name, dims, None)
            return None
        if head in _TYPES:
            typ, rest = self._split_type(toks)
            for name, dims in _parse_entities(rest):
                decl(name, dims, typ)
            return None
        if head == "common":
            raise SyntaxError("%s: COMMON outside common.h not supported" % where)
        if head == "call":
            name = toks[1][1]
            args = []
            if len(toks) > 2:
                inner = toks[3:-1]
                for a in _split_top(inner):
                    if not a:
                        continue
                    args.append(self._parse_actual(a))
            return with_label(("call", name, args, where))
        if head == "if":
            # find matching paren
            depth, j = 0, 1
            for j in range(1, len(toks)):
                if toks[j] == ("op", "("):
                    depth += 1
                elif toks[j] == ("op", ")"):
                    depth -= 1
                    if depth == 0:
                        break
            cond = compile_expr(parse_expr_tokens(toks[2:j]))
            rest = toks[j + 1:]
            if rest == [("id", "then")]:
                branches, else_block = [], None
                while True:
                    block, term, ttoks, _ = parse_block(("elseif", "else", "endif"))
                    branches.append((cond, block))
                    pos[0] += 1
                    if term == "endif":
                        break
                    if term == "elseif":
                        k0 = 2 if ttoks[0][1] == "else" else 1
                        depth = 0
                        for j in range(k0, len(ttoks)):
                            if ttoks[j] == ("op", "("):
                                depth += 1
                            elif ttoks[j] == ("op", ")"):
                                depth -= 1
                                if depth == 0:
                                    break
                        cond = compile_expr(parse_expr_tokens(ttoks[k0 + 1:j]))
                        continue
                    if term == "else":
                        else_block, term2, _, _ = parse_block(("endif",))
                        pos[0] += 1
                        break
                    raise SyntaxError("%s: unterminated IF" % where)
                return with_label(("if", branches, else_block, where))
            inner = self._parse_stmt(unit, rest, text, no, parse_block, pos, decl, None)
            return with_label(("if", [(cond, [inner] if inner is not None else [])], None, where))
        if head == "select":
            # SELECT CASE (expr) / CASE (v1, v2) / CASE DEFAULT / END SELECT  ->  an IF chain
            sel = toks[2:]          # "( expr )"
            _, term, ttoks, _ = parse_block(("case", "endselect"))
            branches, else_block = [], None
            while term == "case":
                pos[0] += 1
                if len(ttoks) > 1 and ttoks[1] == ("id", "default"):
                    else_block, term, ttoks, _ = parse_block(("case", "endselect"))
                    continue
                ct = []
                for v in _split_top(ttoks[2:-1]):
                    if ct:
                        ct.append(("op", ".or."))
                    ct += sel + [("op", ".eq."), ("op", "(")] + v + [("op", ")")]
                cond = compile_expr(parse_expr_tokens(ct))
                block, term, ttoks, _ = parse_block(("case", "endselect"))
                branches.append((cond, block))
            if term != "endselect":
                raise SyntaxError("%s: unterminated SELECT CASE" % where)
            pos[0] += 1
            return with_label(("if", branches, else_block, where))
        if head == "do":
            j = 1
            dolabel = None
            if toks[1][0] == "num":
                dolabel = int(toks[1][1])
                j = 2
            if j >= len(toks):  # bare DO
                raise SyntaxError("%s: DO forever not supported" % where)
            if toks[j] == ("id", "while"):
                cond = compile_expr(parse_expr_tokens(toks[j + 1:]))
                block, term, _, _ = parse_block(("enddo",))
                pos[0] += 1
                return with_label(("dowhile", cond, block, where))
            var = toks[j][1]
            parts = _split_top(toks[j + 2:])
            codes = [compile_expr(parse_expr_tokens(p)) for p in parts]
            if dolabel is None:
                block, term, _, _ = parse_block(("enddo",))
                pos[0] += 1
            else:
                block, term, _, _ = parse_block((("label", dolabel), "enddo"))
                pos[0] += 1
            return with_label(("do", var, codes, block, where))
        if head == "where":
            depth, j = 0, 1
            for j in range(1, len(toks)):
                if toks[j] == ("op", "("):
                    depth += 1
                elif toks[j] == ("op", ")"):
                    depth -= 1
                    if depth == 0:
                        break
            mask = compile_expr(parse_expr_tokens(toks[2:j]))
            rest = toks[j + 1:]
            if rest:
                inner = self._parse_assign(rest, where)
                return with_label(("where", [(mask, [inner])], where))
            branches = []
            while True:
                block, term, ttoks, _ = parse_block(("elsewhere", "endwhere"))
                branches.append((mask, block))
                pos[0] += 1
                if term == "endwhere":
                    break
                if term == "elsewhere":
                    mask = compile_expr(parse_expr_tokens(ttoks[2:-1])) if len(ttoks) > 1 else None
                    continue
                raise SyntaxError("%s: unterminated WHERE" % where)
            return with_label(("where", branches, where))
        if head == "return":
            return with_label(("return",))
        if head == "cycle":
            return with_label(("cycle",))
        if head == "exit":
            return with_label(("exit",))
        if head == "continue":
            return with_label(None)
        if head == "stop":
            return with_label(("stop", where))
        if head in ("goto", "go"):
            target = int(toks[-1][1])
            return with_label(("goto", target))
        if head == "allocate" and ("op", "%") in toks:
            inner = toks[2:-1]
            k = inner.index(("op", "%"))
            base, field = inner[0][1], inner[k + 1][1]
            idx = compile_expr(parse_expr_tokens(inner[2:k - 1]))
            dims = _parse_entities([("id", "_c")] + inner[k + 2:])[0][1]
            return with_label(("allocate_comp", base, idx, field, dims, where))
        if head == "allocate":
            ents = _parse_entities(toks[2:-1])
            return with_label(("allocate", ents, where))
        if head == "deallocate":
            return with_label(None)
        if head in ("end", "contains"):
            return None
        raise SyntaxError("%s: cannot parse statement %r" % (where, text))

    @staticmethod
    def _is_assignment(toks):
        depth = 0
        for k, t in enumerate(toks):
            if t == ("op", "("):
                depth += 1
            elif t == ("op", ")"):
                depth -= 1
            elif depth == 0 and t == ("op", "="):
                if toks[0][1] == "do" and any(x == ("op", ",") for x in _top_level(toks[k + 1:])):
                    return False
                if toks[0][1] in ("if", "where") and toks[1] == ("op", "("):
                    return False
                return True
            elif depth == 0 and k > 0 and t[0] == "id" and toks[k - 1] == ("op", ")") and toks[0][1] in ("if", "where"):
                return False
        return False

    def _parse_assign(self, toks, where):
        depth = 0
        for k, t in enumerate(toks):
            if t == ("op", "("):
                depth += 1
            elif t == ("op", ")"):
                depth -= 1
            elif depth == 0 and t == ("op", "="):
                break
        lhs, rhs = toks[:k], toks[k + 1:]
        name = lhs[0][1]
        rhs_code = compile_expr(parse_expr_tokens(rhs))
        if len(lhs) == 1:
            return ("assign", name, None, rhs_code, where)
        if ("op", "%") in lhs:
            # derived-type component target: evaluate the object expression
            tgt = parse_expr_tokens(lhs)
            return ("assign_obj", compile_expr(tgt), rhs_code, where)
        p = ExprParser(lhs, 2)
        args = p.arglist()
        idx = compile_expr("(%s)" % "".join(a + "," for a in args))
        return ("assign", name, idx, rhs_code, where)

    def _parse_actual(self, a):
        """actual argument -> (kind, ...): ('name', n) | ('elem', n, idxcode) | ('expr', code)"""
        if len(a) == 1 and a[0][0] == "id":
            return ("name", a[0][1], compile_expr(a[0][1]))
        if a[0][0] == "id" and len(a) > 3 and a[1] == ("op", "(") and a[-1] == ("op", ")") and ("op", "%") not in a:
            depth, closed_early = 0, False
            for k in range(1, len(a)):
                if a[k] == ("op", "("):
                    depth += 1
                elif a[k] == ("op", ")"):
                    depth -= 1
                    if depth == 0 and k != len(a) - 1:
                        closed_early = True
                        break
            if not closed_early:
                p = ExprParser(a, 2)
                args = p.arglist()
                if not any(x.startswith("_sl(") for x in args):
                    idx = compile_expr("(%s)" % "".join(x + "," for x in args))
                    return ("elem", a[0][1], idx, compile_expr(parse_expr_tokens(a)))
        return ("expr", compile_expr(parse_expr_tokens(a)))

    # ---- execution --------------------------------------------------------
    def call(self, name, *actuals):
        """call a unit from python; numpy arrays are passed by reference.
        Returns the dict of final dummy-argument values (for scalar outputs)."""
        name = name.lower()
        if name in self.stubs:
            return self.stubs[name](self, *actuals)
        unit = self.units[name]
        L = {}
        for k, dummy in enumerate(unit.args):
            if k < len(actuals):
                L[dummy] = actuals[k]
        for dummy in unit.args:       # scalar dummy fed from a sequence-associated element view
            d = unit.decls.get(dummy)
            if (d is None or d.dims is None) and isinstance(L.get(dummy), np.ndarray) and L[dummy].ndim == 1 \
                    and unit.name in FUNCTION_UNITS:
                v = L[dummy][0]
                L[dummy] = int(v) if L[dummy].dtype.kind in "iu" else float(v)
        env = _Env(self, L, unit)
        copyback = []
        for pname, code in unit.params:
            L[pname] = eval(compile_expr(code), HELPERS, env)
        # declared arrays
        for dname in unit.order:
            d = unit.decls[dname]
            if d.dims is None:
                # a character scalar starts out blank (its content is undefined in Fortran; file names built by
                # formatted internal writes, which are not interpreted, only ever reach python stubs)
                if d.typ == "character" and dname not in unit.args and dname not in L \
                        and dname not in self.M and not (unit.uses_common and dname in self.G):
                    L[dname] = ""
                continue
            if d.alloc:
                continue
            if dname in unit.args:
                if dname not in L:
                    continue
                act = L[dname]
                if act is None:      # absent optional / undefined actual: stays unbound
                    del L[dname]
                    continue
                if isinstance(act, (list, tuple)):
                    continue
                if not isinstance(act, np.ndarray):
                    # scalar actual for an array dummy (asimfg.f passes `EGmassd = one` for EGmass): only the
                    # first element exists; touching anything else raises IndexError
                    L[dname] = np.array([act], dtype=np.float64 if isinstance(act, float) else np.int64)
                    continue
                shape, lbs = self._eval_dims(d.dims, env, total=act.size)
                if tuple(shape) != act.shape:
                    n = int(np.prod(shape))
                    copied = False
                    if act.flags.f_contiguous or act.ndim == 1:
                        flat = act.reshape(-1, order="F")
                    else:
                        # non-contiguous section with a different dummy shape: copy-in / copy-out
                        flat = act.flatten(order="F")
                        copyback.append((act, flat))
                        copied = True
                    if n > flat.size and len(shape) > 1:
                        # declared larger than the actual (e.g. bc3lhs.f BC(nshg,11)): only the part
                        # that exists is addressable; an access beyond it raises IndexError
                        lead = int(np.prod(shape[:-1]))
                        shape[-1] = flat.size // lead if lead else 0
                        n = int(np.prod(shape))
                    if n > flat.size:
                        raise ValueError("%s: dummy %s%r larger than actual (%d elements)"
                                         % (unit.name, dname, tuple(shape), flat.size))
                    v = flat[:n].reshape(shape, order="F")
                    if not copied and not np.shares_memory(v, act) and n > 0:
                        raise ValueError("%s: dummy %s lost aliasing" % (unit.name, dname))
                    L[dname] = v
                if any(l != 1 for l in lbs):
                    env.local_lb[dname] = tuple(lbs)
            elif unit.uses_common and dname in self.G:
                continue
            else:
                shape, lbs = self._eval_dims(d.dims, env, total=None)
                typ = d.typ or ("integer" if dname[0] in "ijklmn" else "real")
                if typ == "integer":
                    L[dname] = np.zeros(shape, dtype=np.int64, order="F")
                elif typ == "logical":
                    L[dname] = np.zeros(shape, dtype=np.bool_, order="F")
                else:
                    L[dname] = np.full(shape, np.nan if self.nan_locals else 0.0, dtype=np.float64, order="F")
                if any(l != 1 for l in lbs):
                    env.local_lb[dname] = tuple(lbs)
        try:
            self._exec_block(unit.body, env)
        except _Return:
            pass
        for act, flat in copyback:
            act[...] = flat.reshape(act.shape, order="F")
        return L

    def _eval_dims(self, dims, env, total):
        shape, lbs = [], []
        for lo, hi in dims:
            l = 1 if lo is None else int(eval(compile_expr(lo), HELPERS, env))
            lbs.append(l)
            if hi in ("*", ":"):
                shape.append(None)
            else:
                shape.append(int(eval(compile_expr(hi), HELPERS, env)) - l + 1)
        if None in shape:
            known = int(np.prod([s for s in shape if s is not None])) if len(shape) > 1 else 1
            shape[shape.index(None)] = (total // known) if (total is not None and known > 0) else 0
        shape = [max(s, 0) for s in shape]
        return shape, lbs

    def _exec_block(self, block, env):
        i = 0
        n = len(block)
        while i < n:
            try:
                self._exec(block[i], env)
            except _Goto as g:
                j = self._find_label(block, g.label)
                if j is None:
                    raise
                i = j
                continue
            i += 1

    @staticmethod
    def _find_label(block, label):
        for j, s in enumerate(block):
            if s[0] == "label" and s[1] == label:
                return j
            if s[0] == "seq" and s[1] and s[1][0][0] == "label" and s[1][0][1] == label:
                return j
        return None

    def _exec(self, s, env):
        kind = s[0]
        if kind == "assign":
            _, name, idx, rhs, where = s
            try:
                val = eval(rhs, HELPERS, env)
                self._assign(env, name, None if idx is None else eval(idx, HELPERS, env), val)
            except (_Return, _Cycle, _Exit, _Goto):
                raise
            except Exception as e:
                raise type(e)("%s [%s]" % (e, where)) from e
            if self.trace:
                self.trace(where, name, env)
        elif kind == "if":
            for cond, block in s[1]:
                try:
                    c = eval(cond, HELPERS, env)
                except Exception as e:
                    raise type(e)("%s [%s]" % (e, s[3])) from e
                if np.ndim(c) != 0:
                    raise TypeError("array-valued IF condition [%s]" % s[3])
                if c:
                    self._exec_block(block, env)
                    return
            if s[2] is not None:
                self._exec_block(s[2], env)
        elif kind == "do":
            _, var, codes, block, where = s
            lo = int(eval(codes[0], HELPERS, env))
            hi = int(eval(codes[1], HELPERS, env))
            st = int(eval(codes[2], HELPERS, env)) if len(codes) > 2 else 1
            i = lo
            while (st > 0 and i <= hi) or (st < 0 and i >= hi):
                self._assign(env, var, None, i)
                try:
                    self._exec_block(block, env)
                except _Cycle:
                    pass
                except _Exit:
                    break
                i += st
            else:
                self._assign(env, var, None, i)
        elif kind == "dowhile":
            while eval(s[1], HELPERS, env):
                try:
                    self._exec_block(s[2], env)
                except _Cycle:
                    continue
                except _Exit:
                    break
        elif kind == "call":
            self._exec_call(s, env)
        elif kind == "where":
            done = None
            with np.errstate(all="ignore"):
                for mask_code, block in s[1]:
                    if mask_code is None:
                        m = np.logical_not(done)
                    else:
                        m = np.asarray(eval(mask_code, HELPERS, env), dtype=bool)
                        if done is not None:
                            m = np.logical_and(m, np.logical_not(done))
                    done = m if done is None else np.logical_or(done, m)
                    for st in block:
                        if st[0] == "seq":
                            sub = [x for x in st[1] if x[0] != "label"]
                        else:
                            sub = [st]
                        for a in sub:
                            if a[0] != "assign":
                                raise SyntaxError("non-assignment inside WHERE [%s]" % s[2])
                            _, name, idx, rhs, where = a
                            val = eval(rhs, HELPERS, env)
                            tgt = env[name]
                            if idx is not None:
                                tgt = tgt[_conv_index(eval(idx, HELPERS, env), env.lbounds(name))]
                            tgt[m] = np.broadcast_to(val, tgt.shape)[m]
        elif kind == "seq":
            for x in s[1]:
                self._exec(x, env)
        elif kind == "label":
            pass
        elif kind == "return":
            raise _Return()
        elif kind == "cycle":
            raise _Cycle()
        elif kind == "exit":
            raise _Exit()
        elif kind == "goto":
            raise _Goto(s[1])
        elif kind == "stop":
            raise RuntimeError("f77np: STOP at %s" % s[1])
        elif kind == "allocate":
            for name, dims in s[1]:
                shape, lbs = self._eval_dims(dims, env, None)
                d = env.unit.decls.get(name)
                typ = (d.typ if d and d.typ else None) or ("integer" if name[0] in "ijklmn" else "real")
                dt = np.int64 if typ == "integer" else np.float64
                env.L[name] = np.full(shape, np.nan if (dt == np.float64 and self.nan_locals) else 0, dtype=dt, order="F")
        elif kind == "allocate_comp":
            # module array of derived-type pointers (pointer.f:42-47), kept by the driver as a python list in M;
            # M["_comp_dtype"][base] says integer / real (the module declaration is not parsed)
            _, base, idx, field, dims, where = s
            lst = self.M[base]
            i = int(eval(idx, HELPERS, env))
            while len(lst) < i:
                lst.append(CompPtr())
            shape, lbs = self._eval_dims(dims, env, None)
            dt = self.M.get("_comp_dtype", {}).get(base, np.float64)
            setattr(lst[i - 1], field, np.full(shape, np.nan if dt == np.float64 else 0, dtype=dt, order="F"))
        elif kind == "assign_obj":
            tgt = eval(s[1], HELPERS, env)
            tgt[...] = eval(s[2], HELPERS, env)
        else:
            raise RuntimeError("f77np: unknown statement kind %r" % (kind,))

    def _assign(self, env, name, idx, val):
        L = env.L
        if idx is not None:
            arr = env[name]
            if isinstance(arr, list):
                arr[int(idx[0]) - 1] = val
                return
            arr[_conv_index(idx, env.lbounds(name))] = val
            return
        if name in L:
            cur, store = L[name], L
        elif env.unit is not None and env.unit.uses_common and name in self.G:
            cur, store = self.G[name], self.G
        elif name in self.M:
            cur, store = self.M[name], self.M
        else:
            cur, store = None, L
        if isinstance(cur, np.ndarray):
            cur[...] = val
            return
        if isinstance(val, np.ndarray) and val.ndim > 0:
            raise TypeError("f77np: array value assigned to scalar %r" % name)
        d = env.unit.decls.get(name) if env.unit is not None else None
        if store is L:
            typ = d.typ if d is not None else None
        else:
            typ = self.gtypes.get(name)
        if typ is None:
            typ = "integer" if name[0] in "ijklmn" else "real"
        if isinstance(val, str) or typ == "character":
            store[name] = val
        elif typ == "logical":
            store[name] = bool(val)
        elif typ == "integer":
            store[name] = int(val)  # truncation toward zero
        else:
            store[name] = float(val)

    def _exec_call(self, s, env):
        _, name, args, where = s
        actuals, writeback, elem_out = [], [], []
        callee = self.units.get(name)
        try:
            for k, a in enumerate(args):
                if a[0] == "name":
                    try:
                        v = eval(a[2], HELPERS, env)
                    except NameError:
                        v = None  # undefined local passed as output
                    actuals.append(v)
                    if not isinstance(v, np.ndarray):
                        writeback.append((k, a[1]))
                elif a[0] == "elem":
                    arr = env[a[1]] if a[1] in env else None
                    wants_array = False
                    if callee is not None and k < len(callee.args):
                        d = callee.decls.get(callee.args[k])
                        wants_array = d is not None and d.dims is not None
                    elif name in self.stubs:
                        # a python stub may ask for sequence association too (MPI buffers: `global(isgbeg,1)`)
                        wants_array = k in getattr(self.stubs[name], "array_args", ())
                    if isinstance(arr, np.ndarray) and wants_array:
                        idx = _conv_index(eval(a[2], HELPERS, env), env.lbounds(a[1]))
                        off = int(np.ravel_multi_index(idx, arr.shape, order="F"))
                        if not arr.flags.f_contiguous:
                            raise ValueError("sequence association on a non-contiguous array")
                        actuals.append(arr.reshape(-1, order="F")[off:])
                    else:
                        actuals.append(eval(a[3], HELPERS, env))
                        if name in self.stubs and isinstance(arr, np.ndarray):
                            # an array element as an output argument of a python stub (sevsegtype(itask,1))
                            elem_out.append((k, a[1], eval(a[2], HELPERS, env)))
                else:
                    actuals.append(eval(a[1], HELPERS, env))
            if name in self.stubs:
                out = self.stubs[name](self, *actuals)
                if isinstance(out, dict):
                    for k, nm in writeback:
                        if k in out:
                            self._assign(env, nm, None, out[k])
                    for k, nm, idx in elem_out:
                        if k in out:
                            self._assign(env, nm, idx, out[k])
                return
            if callee is None:
                raise NameError("f77np: subroutine %r is not loaded" % name)
            Lc = self.call(name, *actuals)
        except (_Return, _Cycle, _Exit, _Goto):
            raise
        except Exception as e:
            if "[call " not in str(e)[-80:]:
                raise type(e)("%s [call %s at %s]" % (e, name, where)) from e
            raise
        for k, nm in writeback:
            if k < len(callee.args):
                dv = Lc.get(callee.args[k])
                if dv is not None and not isinstance(dv, np.ndarray):
                    if actuals[k] is None or dv != actuals[k] or type(dv) is not type(actuals[k]):
                        self._assign(env, nm, None, dv)


def _top_level(toks):
    depth = 0
    for t in toks:
        if t == ("op", "("):
            depth += 1
        elif t == ("op", ")"):
            depth -= 1
        elif depth == 0:
            yield t


class _Env(dict):
    """locals mapping for eval(): locals -> COMMON/PARAMETER -> module variables"""

    def __init__(self, prog, L, unit):
        super().__init__()
        self.prog, self.L, self.unit = prog, L, unit
        self.local_lb = {}

    def __getitem__(self, k):
        if k == "_E":
            return self
        L = self.L
        if k in L:
            return L[k]
        G = self.prog.G
        if k in G:
            return G[k]
        M = self.prog.M
        if k in M:
            return M[k]
        raise KeyError(k)

    def __contains__(self, k):
        return k in self.L or k in self.prog.G or k in self.prog.M

    def lbounds(self, name):
        if name in self.local_lb:
            return self.local_lb[name]
        if name not in self.L:
            return self.prog.gl_lbounds.get(name)
        return None

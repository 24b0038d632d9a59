"""Script host for the CLI batch mode: an interpreter for the part of Rhai that PaintFE scripts use, wired to the
host API the reference registers (src/ops/scripting.rs:323-1481: canvas, pixel, transform, effect, utility and
selection APIs).

This is CALLER-side code (SURVEY section 2.1 keeps the Rhai host as the caller of the hot path): every
`apply_*`, flip / rotate / resize and fill call goes to the device through paintfe_b200.script.bindings, exactly as
the straight-line runner did.  What the interpreter adds is the host language around those calls - `let`, `if`,
`while` / `for` / `loop`, `fn`, arrays, strings with `${}` interpolation, closures - so that scripts written for the
reference run unchanged.

Closures passed to for_each_pixel / for_region / map_channels are user code which the reference interprets once per
pixel on the CPU (scripting.rs:437-616).  They are user code here too: the body is evaluated ONCE over whole-image
int64 arrays of x, y, r, g, b, a (a data-dependent `if` evaluates both arms and selects), with a per-pixel
evaluation of the same body as the fallback for bodies that cannot be evaluated that way (loops on pixel values, side
effects).  No kernel of the hot path is replaced by this; nothing here imports oracle/.

Language covered: integer (i64) / float (f64) / bool / string / array / unit values; + - * / % ** and the bit
operators, comparisons, && || !; compound assignment; `if` as statement and expression, `while`, `loop`,
`for v in a..b | a..=b | range(a, b[, step]) | array`, `break` / `continue` / `return`; `fn`; `|args| expr`
closures; indexing; the methods len / push / pop / to_int / to_float / abs / min / max / contains / to_string.
Not covered: object maps, switch, modules, try / catch, string methods beyond len / to_string.
"""
from __future__ import annotations

import math
import re
from typing import Any, Callable, Dict, List, Optional

import numpy as np


class ScriptError(ValueError):
    """ScriptError of scripting.rs:90: message plus the source line when it is known."""

    def __init__(self, message: str, line: Optional[int] = None):
        super().__init__(message if line is None else f"Line {line}: {message}")
        self.message, self.line = message, line


class _NotVectorisable(Exception):
    pass


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    def __init__(self, value):
        self.value = value


_TOKEN = re.compile(r"""
   (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
 | (?P<float>\d[\d_]*\.\d[\d_]*(?:[eE][+-]?\d+)?|\d[\d_]*[eE][+-]?\d+)
 | (?P<int>0x[0-9a-fA-F_]+|\d[\d_]*)
 | (?P<str>"(?:\\.|[^"\\])*")
 | (?P<tpl>`(?:\\.|[^`\\])*`)
 | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
 | (?P<op>\.\.=|\*\*=|<<=|>>=|\.\.|\*\*|==|!=|<=|>=|&&|\|\||\+=|-=|\*=|/=|%=|&=|\|=|\^=|<<|>>|[-+*/%<>=!&|^(){}\[\],;.:])
""", re.X | re.S)

_KEYWORDS = {"let", "const", "if", "else", "while", "loop", "for", "in", "break", "continue", "return", "fn", "true", "false"}
_ESC = {"n": "\n", "t": "\t", "r": "\r", "\\": "\\", '"': '"', "`": "`", "0": "\0", "$": "$"}


def _unescape(s: str) -> str:
    return re.sub(r"\\(.)", lambda m: _ESC.get(m.group(1), m.group(1)), s)


def tokenize(src: str):
    toks, pos, line = [], 0, 1
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise ScriptError(f"Syntax error: unexpected character {src[pos]!r}", line)
        kind, text = m.lastgroup, m.group()
        if kind != "ws":
            toks.append((kind, text, line))
        line += text.count("\n")
        pos = m.end()
    toks.append(("eof", "", line))
    return toks


# Binary operator precedences (Rhai's table, lowest first)
_PREC = {"||": 30, "|": 30, "^": 30, "&&": 60, "&": 60, "==": 90, "!=": 90, "<": 110, "<=": 110, ">": 110, ">=": 110,
         "..": 120, "..=": 120, "+": 150, "-": 150, "*": 180, "/": 180, "%": 180, "**": 190, "<<": 210, ">>": 210}
_ASSIGN = {"=", "+=", "-=", "*=", "/=", "%=", "**=", "<<=", ">>=", "&=", "|=", "^="}


class Parser:
    def __init__(self, src: str):
        self.t = tokenize(src)
        self.i = 0

    # -- token helpers
    def peek(self, k=0):
        return self.t[min(self.i + k, len(self.t) - 1)]

    def at(self, text):
        k, t, _ = self.peek()
        return t == text and k in ("op", "id")

    def take(self, text=None):
        tok = self.peek()
        if text is not None and tok[1] != text:
            found = tok[1] or "end of script"
            raise ScriptError(f"Syntax error: expecting '{text}', found '{found}'", tok[2])
        self.i += 1
        return tok

    def ident(self):
        k, t, ln = self.take()
        if k != "id" or t in _KEYWORDS:
            raise ScriptError(f"Syntax error: expecting a name, found '{t}'", ln)
        return t

    # -- grammar
    def program(self):
        body = []
        while self.peek()[0] != "eof":
            body.append(self.statement())
        return body

    def block(self):
        self.take("{")
        body = []
        while not self.at("}"):
            if self.peek()[0] == "eof":
                raise ScriptError("Syntax error: expecting '}' to close the block", self.peek()[2])
            body.append(self.statement())
        self.take("}")
        return ("block", body)

    def end_stmt(self, block_like=False):
        """';' terminates a statement; it is optional after a block-like statement and before '}'."""
        if self.at(";"):
            self.take()
            return True
        if block_like or self.at("}") or self.peek()[0] == "eof":
            return False
        tok = self.peek()
        raise ScriptError(f"Syntax error: expecting ';' to terminate this statement, found '{tok[1]}'", tok[2])

    def statement(self):
        k, t, ln = self.peek()
        if k == "op" and t == ";":
            self.take()
            return ("expr", ("unit",), True, ln)
        if k == "id" and t in ("let", "const"):
            self.take()
            name = self.ident()
            init = ("unit",)
            if self.at("="):
                self.take()
                init = self.expr()
            self.end_stmt()
            return ("let", name, init, ln)
        if k == "id" and t == "fn":
            self.take()
            name = self.ident()
            self.take("(")
            params = []
            while not self.at(")"):
                params.append(self.ident())
                if not self.at(")"):
                    self.take(",")
            self.take(")")
            return ("fn", name, params, self.block(), ln)
        if k == "id" and t == "while":
            self.take()
            cond = self.expr()
            body = self.block()
            self.end_stmt(True)
            return ("while", cond, body, ln)
        if k == "id" and t == "loop":
            self.take()
            body = self.block()
            self.end_stmt(True)
            return ("while", ("lit", True), body, ln)
        if k == "id" and t == "for":
            self.take()
            var = self.ident()
            self.take("in")
            it = self.expr()
            body = self.block()
            self.end_stmt(True)
            return ("for", var, it, body, ln)
        if k == "id" and t in ("break", "continue"):
            self.take()
            self.end_stmt()
            return (t, ln)
        if k == "id" and t == "return":
            self.take()
            val = ("unit",) if (self.at(";") or self.at("}")) else self.expr()
            self.end_stmt()
            return ("return", val, ln)
        if (k == "id" and t == "if") or (k == "op" and t == "{"):
            e = self.if_expr() if t == "if" else self.block()
            return ("expr", e, self.end_stmt(True), ln)
        e = self.expr()
        if self.peek()[0] == "op" and self.peek()[1] in _ASSIGN:
            op = self.take()[1]
            if e[0] not in ("var", "index"):
                raise ScriptError("Syntax error: cannot assign to this expression", ln)
            rhs = self.expr()
            self.end_stmt()
            return ("assign", e, op, rhs, ln)
        return ("expr", e, self.end_stmt(), ln)

    def expr(self, min_prec=0):
        lhs = self.unary()
        while True:
            k, t, ln = self.peek()
            if k != "op" or t not in _PREC or _PREC[t] < min_prec:
                return lhs
            self.take()
            # ** binds to the right, the rest to the left
            rhs = self.expr(_PREC[t] if t == "**" else _PREC[t] + 1)
            lhs = ("range", lhs, rhs, t == "..=") if t in ("..", "..=") else ("bin", t, lhs, rhs, ln)

    def unary(self):
        k, t, ln = self.peek()
        if k == "op" and t in ("-", "+", "!"):
            self.take()
            operand = self.unary()
            if t == "-" and operand[0] == "lit" and isinstance(operand[1], (int, float)) and not isinstance(operand[1], bool):
                return ("lit", -operand[1])
            return operand if t == "+" else ("un", t, operand, ln)
        return self.postfix(self.primary())

    def args(self):
        self.take("(")
        out = []
        while not self.at(")"):
            out.append(self.expr())
            if not self.at(")"):
                self.take(",")
        self.take(")")
        return out

    def postfix(self, e):
        while True:
            if self.at("["):
                ln = self.take()[2]
                idx = self.expr()
                self.take("]")
                e = ("index", e, idx, ln)
            elif self.at(".") and self.peek(1)[0] == "id":
                ln = self.take()[2]
                name = self.ident()
                e = ("method", e, name, self.args() if self.at("(") else None, ln)
            else:
                return e

    def closure(self):
        ln = self.peek()[2]
        params = []
        if self.at("||"):
            self.take()
        else:
            self.take("|")
            while not self.at("|"):
                params.append(self.ident())
                if not self.at("|"):
                    self.take(",")
            self.take("|")
        body = self.block() if self.at("{") else self.expr()
        return ("closure", params, body, ln)

    def primary(self):
        k, t, ln = self.peek()
        if k == "int":
            self.take()
            return ("lit", int(t.replace("_", ""), 0))
        if k == "float":
            self.take()
            return ("lit", float(t.replace("_", "")))
        if k == "str":
            self.take()
            return ("lit", _unescape(t[1:-1]))
        if k == "tpl":
            self.take()
            return self.template(t[1:-1], ln)
        if k == "op":
            if t == "(":
                self.take()
                if self.at(")"):
                    self.take()
                    return ("unit",)
                e = self.expr()
                self.take(")")
                return e
            if t == "[":
                self.take()
                items = []
                while not self.at("]"):
                    items.append(self.expr())
                    if not self.at("]"):
                        self.take(",")
                self.take("]")
                return ("arr", items)
            if t == "{":
                return self.block()
            if t in ("|", "||"):
                return self.closure()
        if k == "id":
            if t in ("true", "false"):
                self.take()
                return ("lit", t == "true")
            if t == "if":
                return self.if_expr()
            if t not in _KEYWORDS:
                self.take()
                if self.at("("):
                    return ("call", t, self.args(), ln)
                return ("var", t, ln)
        raise ScriptError(f"Syntax error: unexpected '{t or 'end of script'}'", ln)

    def if_expr(self):
        ln = self.take("if")[2]
        cond = self.expr()
        then = self.block()
        other = None
        if self.at("else"):
            self.take()
            other = self.if_expr() if self.at("if") else self.block()
        return ("if", cond, then, other, ln)

    def template(self, body: str, ln: int):
        parts, pos = [], 0
        while pos < len(body):
            j = body.find("${", pos)
            if j < 0:
                parts.append(("lit", _unescape(body[pos:])))
                break
            if j > pos:
                parts.append(("lit", _unescape(body[pos:j])))
            depth, e = 1, j + 2
            while e < len(body) and depth:
                depth += {"{": 1, "}": -1}.get(body[e], 0)
                e += 1
            if depth:
                raise ScriptError("Syntax error: unterminated ${ in string", ln)
            sub = Parser(body[j + 2:e - 1])
            parts.append(sub.expr())
            if sub.peek()[0] != "eof":
                raise ScriptError("Syntax error: unexpected text inside ${ }", ln)
            pos = e
        return ("tpl", parts)


class Closure:
    def __init__(self, params, body, scopes):
        self.params, self.body, self.scopes = params, body, scopes


def _is_vec(v) -> bool:
    return isinstance(v, np.ndarray)


def _is_int(v) -> bool:
    if _is_vec(v):
        return v.dtype.kind == "i"
    return isinstance(v, int) and not isinstance(v, bool)


def _is_num(v) -> bool:
    if _is_vec(v):
        return v.dtype.kind in "if"
    return isinstance(v, (int, float)) and not isinstance(v, bool)


def to_text(v) -> str:
    """Rhai's Display of a value (what print and ${} produce)."""
    if v is None:
        return ""
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, float):
        if math.isnan(v):
            return "NaN"
        if math.isinf(v):
            return "inf" if v > 0 else "-inf"
        return repr(v) if v != int(v) or abs(v) >= 1e16 else f"{v:.1f}"
    if isinstance(v, list):
        return "[" + ", ".join(f'"{x}"' if isinstance(x, str) else to_text(x) for x in v) + "]"
    if isinstance(v, Closure):
        return "Fn(anonymous)"
    return str(v)


_I64_MIN, _I64_MAX = -(1 << 63), (1 << 63) - 1


def _checked_i64(op, a, b, r, line=None):
    """Rhai's default (checked) integer arithmetic: a result outside i64 is an error, not a bigger integer.
    Whole-image evaluation works on int64 arrays, where an overflow would wrap silently: operands large enough to get
    there send the closure to the per-pixel path, which checks every operation."""
    if _is_vec(r):
        big = 1 << (31 if op in ("*", "**") else 62)
        for v in (a, b):
            if (_is_vec(v) and v.size and int(np.abs(v).max()) >= big) or (not _is_vec(v) and abs(int(v)) >= big):
                raise _NotVectorisable()
        if op == "**":
            with np.errstate(over="ignore"):
                if float(np.max(np.power(np.abs(np.asarray(a, np.float64)), np.asarray(b, np.float64)), initial=0.0)) >= 2.0 ** 62:
                    raise _NotVectorisable()
        return r
    if not _I64_MIN <= r <= _I64_MAX:
        raise ScriptError("Arithmetic overflow" if op != "**" else "Number raised to power overflows", line)
    return r


def _trunc_div(a, b):
    if _is_vec(a) or _is_vec(b):
        b = np.asarray(b)
        if (b == 0).any():
            raise _NotVectorisable()
        q = np.abs(a) // np.abs(b)
        return np.where((np.asarray(a) < 0) != (b < 0), -q, q)
    if b == 0:
        raise ScriptError("Division by zero")
    q = abs(a) // abs(b)
    return -q if (a < 0) != (b < 0) else q


def _trunc_rem(a, b):
    if _is_vec(a) or _is_vec(b):
        return np.asarray(a) - _trunc_div(a, b) * b
    if b == 0:
        raise ScriptError("Modulo division by zero")
    return a - _trunc_div(a, b) * b


def _binary(op, a, b, line=None):
    if op == "+" and (isinstance(a, str) or isinstance(b, str)):
        return to_text(a) + to_text(b)
    if op == "+" and isinstance(a, list) and isinstance(b, list):
        return a + b
    if op in ("==", "!="):
        if _is_vec(a) or _is_vec(b):
            return (a == b) if op == "==" else (a != b)
        same = isinstance(a, bool) == isinstance(b, bool) and type(a) in (type(b), int, float) and a == b
        return same if op == "==" else not same
    if isinstance(a, bool) and isinstance(b, bool) or (_is_vec(a) and a.dtype == bool) or (_is_vec(b) and b.dtype == bool):
        if op in ("&", "|", "^"):
            f = {"&": np.logical_and, "|": np.logical_or, "^": np.logical_xor}[op]
            r = f(a, b)
            return r if _is_vec(r) else bool(r)
    if isinstance(a, str) and isinstance(b, str) and op in ("<", "<=", ">", ">="):
        return {"<": a < b, "<=": a <= b, ">": a > b, ">=": a >= b}[op]
    if not (_is_num(a) and _is_num(b)):
        raise ScriptError(f"Function not found: {op} ({_type_name(a)}, {_type_name(b)})", line)
    ints = _is_int(a) and _is_int(b)
    if op == "+":
        return _checked_i64(op, a, b, a + b, line) if ints else a + b
    if op == "-":
        return _checked_i64(op, a, b, a - b, line) if ints else a - b
    if op == "*":
        return _checked_i64(op, a, b, a * b, line) if ints else a * b
    if op == "/":
        if ints:
            return _trunc_div(a, b)
        if _is_vec(a) or _is_vec(b):
            with np.errstate(divide="ignore", invalid="ignore"):
                return np.asarray(a, np.float64) / np.asarray(b, np.float64)
        if b == 0:
            return math.copysign(math.inf, a) * math.copysign(1.0, b) if a != 0 else math.nan
        return a / b
    if op == "%":
        if ints:
            return _trunc_rem(a, b)
        return np.fmod(a, b) if (_is_vec(a) or _is_vec(b)) else (math.fmod(a, b) if b != 0 else math.nan)
    if op == "**":
        if ints:
            if _is_vec(a) or _is_vec(b):
                if (np.asarray(b) < 0).any():
                    raise _NotVectorisable()
                return _checked_i64(op, a, b, np.power(a, b), line)
            if b < 0:
                raise ScriptError("Integer raised to a negative power", line)
            if abs(a) > 1 and b > 64:
                raise ScriptError("Number raised to power overflows", line)
            return _checked_i64(op, a, b, a ** b, line)
        if _is_vec(a) or _is_vec(b):
            return np.power(np.asarray(a, np.float64), b)
        try:
            return math.pow(a, b)
        except (OverflowError, ValueError):
            return math.inf
    if op in ("<", "<=", ">", ">="):
        r = {"<": lambda: a < b, "<=": lambda: a <= b, ">": lambda: a > b, ">=": lambda: a >= b}[op]()
        return r if _is_vec(r) else bool(r)
    if op in ("&", "|", "^", "<<", ">>") and ints:
        return {"&": lambda: a & b, "|": lambda: a | b, "^": lambda: a ^ b, "<<": lambda: a << b, ">>": lambda: a >> b}[op]()
    raise ScriptError(f"Function not found: {op} ({_type_name(a)}, {_type_name(b)})", line)


def _type_name(v) -> str:
    if v is None:
        return "()"
    if isinstance(v, bool) or (_is_vec(v) and v.dtype == bool):
        return "bool"
    if _is_int(v):
        return "i64"
    if isinstance(v, float) or _is_vec(v):
        return "f64"
    return {str: "string", list: "array", Closure: "Fn"}.get(type(v), type(v).__name__)


def _as_f(v):
    return np.asarray(v, np.float64) if _is_vec(v) else float(v)


def _round_half_away(v):
    """f64::round: halves away from zero (numpy and Python round to even)."""
    if _is_vec(v):
        return np.copysign(np.floor(np.abs(v) + 0.5), v)
    return math.copysign(math.floor(abs(v) + 0.5), v)


def _to_int(v):
    if _is_vec(v):
        return np.trunc(v).astype(np.int64) if v.dtype.kind == "f" else v.astype(np.int64)
    if isinstance(v, float) and not math.isfinite(v):
        raise ScriptError("Integer overflow: to_int(" + to_text(v) + ")")
    return int(v)


def _vec_fn(scalar: Callable, vector: Callable):
    return lambda *a: vector(*a) if any(_is_vec(x) for x in a) else scalar(*a)


def _rgb_to_hsl(r, g, b):  # scripting.rs:1292
    rf, gf, bf = (min(max(int(c), 0), 255) / 255.0 for c in (r, g, b))
    mx, mn = max(rf, gf, bf), min(rf, gf, bf)
    l = (mx + mn) / 2.0
    if abs(mx - mn) < 1e-10:
        return [0.0, 0.0, l * 100.0]
    d = mx - mn
    s = d / (2.0 - mx - mn) if l > 0.5 else d / (mx + mn)
    if abs(mx - rf) < 1e-10:
        h = (gf - bf) / d + (6.0 if gf < bf else 0.0)
    elif abs(mx - gf) < 1e-10:
        h = (bf - rf) / d + 2.0
    else:
        h = (rf - gf) / d + 4.0
    return [h * 60.0, s * 100.0, l * 100.0]


def _hsl_to_rgb(h, s, l):  # scripting.rs:1330
    s, l = float(s) / 100.0, float(l) / 100.0
    c = (1.0 - abs(2.0 * l - 1.0)) * s
    h2 = float(h) / 60.0
    x = c * (1.0 - abs(math.fmod(h2, 2.0) - 1.0))
    sector = int(max(min(h2, 2147483647.0), -2147483648.0)) if math.isfinite(h2) else 0
    r1, g1, b1 = {0: (c, x, 0.0), 1: (x, c, 0.0), 2: (0.0, c, x), 3: (0.0, x, c), 4: (x, 0.0, c)}.get(sector, (c, 0.0, x))
    m = l - c / 2.0
    return [int(_round_half_away((v + m) * 255.0)) for v in (r1, g1, b1)]


_MATH: Dict[str, Callable] = {
    "clamp": _vec_fn(lambda v, lo, hi: min(max(v, lo), hi), lambda v, lo, hi: np.clip(v, lo, hi)),
    "lerp": lambda a, b, t: _as_f(a) + (_as_f(b) - _as_f(a)) * _as_f(t),
    "distance": _vec_fn(lambda x1, y1, x2, y2: math.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2),
                        lambda x1, y1, x2, y2: np.sqrt((_as_f(x2) - _as_f(x1)) ** 2 + (_as_f(y2) - _as_f(y1)) ** 2)),
    "abs": _vec_fn(abs, np.abs),
    "min": _vec_fn(min, np.minimum),
    "max": _vec_fn(max, np.maximum),
    "floor": _vec_fn(lambda x: float(math.floor(x)), lambda x: np.floor(_as_f(x))),
    "ceil": _vec_fn(lambda x: float(math.ceil(x)), lambda x: np.ceil(_as_f(x))),
    "round": lambda x: _round_half_away(_as_f(x)),
    "sqrt": _vec_fn(lambda x: math.sqrt(x) if x >= 0 else math.nan, lambda x: np.sqrt(_as_f(x))),
    "pow": _vec_fn(lambda x, y: _binary("**", float(x), float(y)), lambda x, y: np.power(_as_f(x), _as_f(y))),
    "sin": _vec_fn(math.sin, np.sin), "cos": _vec_fn(math.cos, np.cos), "tan": _vec_fn(math.tan, np.tan),
    "atan2": _vec_fn(math.atan2, np.arctan2),
    "PI": lambda: math.pi,
    "to_int": _to_int, "to_float": _as_f,
}
for _alias, _base in (("clamp_f", "clamp"), ("abs_i", "abs"), ("min_i", "min"), ("max_i", "max"), ("min_f", "min"), ("max_f", "max")):
    _MATH[_alias] = _MATH[_base]


def _copy_value(v):
    return [_copy_value(x) for x in v] if isinstance(v, list) else v


_PROGRAMS: Dict[str, list] = {}  # source text -> parsed program (the interpreter never mutates a program)


class Interpreter:
    """One script run over one image (the ScriptContext of scripting.rs:262)."""

    MAX_OPERATIONS = 50_000_000  # engine.set_max_operations guard of the reference's worker

    def __init__(self, eng, pixels, mask=None, exact: bool = False, seed: int = 0x2545F4914F6CDD1D):
        from . import script as _script
        self._script = _script
        self.eng, self.img, self.mask = eng, pixels, mask
        self.effects = _script.bindings(eng, exact)
        self.console: List[str] = []
        self.canvas_ops: List[tuple] = []  # CanvasOpRequest log (scripting.rs:42): canvas-wide calls to replay on the other layers
        self.rng_state = (seed & 0xFFFFFFFFFFFFFFFF) or 1
        self.functions: Dict[str, tuple] = {}
        self.scopes: List[Dict[str, Any]] = [{}]
        self.vector_mode = False
        self._outer_scopes = frozenset()  # ids of the scope dicts a whole-image closure evaluation must not write to
        self._outer_lists = frozenset()   # ids of the arrays reachable from them
        self.ops = 0
        self.bulk_evaluations = {"whole_image": 0, "per_pixel": 0}  # how closures of the bulk calls were evaluated
        self._host: Optional[np.ndarray] = None  # host copy of a device image while pixel access is in use
        self._host_dirty = False

    # ---------------------------------------------------------------- image state
    @property
    def width(self) -> int:
        return int(self.img.shape[1])

    @property
    def height(self) -> int:
        return int(self.img.shape[0])

    def host_pixels(self) -> np.ndarray:
        if isinstance(self.img, np.ndarray):
            return self.img
        if self._host is None:
            self._host = self.img.cpu().numpy()
        return self._host

    def touch_host(self):
        if isinstance(self.img, np.ndarray):
            if not self.img.flags.writeable or self._host is None:
                self.img = self._host = np.array(self.img)  # never write into the caller's buffer
        else:
            self.host_pixels()
            self._host_dirty = True

    def device_image(self):
        """The image as the effect entry points take it, with pending set_pixel writes carried over."""
        if not isinstance(self.img, np.ndarray) and self._host_dirty:
            import torch
            self.img = torch.from_numpy(self._host).to(self.img.device)
        self._host_dirty = False
        return self.img

    def set_image(self, img):
        self.img, self._host, self._host_dirty = img, None, False

    def host_mask(self) -> Optional[np.ndarray]:
        m = self.mask
        return m if (m is None or isinstance(m, np.ndarray)) else m.cpu().numpy()

    # ---------------------------------------------------------------- entry points
    def run(self, source: str):
        program = _PROGRAMS.get(source)  # a batch runs one script over many images: parse it once
        if program is None:
            program = Parser(source).program()
            if len(_PROGRAMS) > 64:
                _PROGRAMS.clear()
            _PROGRAMS[source] = program
        for st in program:  # functions are visible before their definition, as in Rhai
            if st[0] == "fn":
                self.functions[st[1]] = (st[2], st[3])
        try:
            self.exec_block(program, new_scope=False)
        except _Return:
            pass
        except (_Break, _Continue):
            raise ScriptError("break / continue outside a loop")
        except RecursionError:
            raise ScriptError("Stack overflow: too many nested function calls")
        return self.device_image()

    # ---------------------------------------------------------------- statements
    def exec_block(self, body, new_scope=True):
        if new_scope:
            self.scopes.append({})
        try:
            value = None
            for st in body:
                value = self.exec(st)
            return value
        finally:
            if new_scope:
                self.scopes.pop()

    def lookup(self, name, line=None):
        for sc in reversed(self.scopes):
            if name in sc:
                return sc
        raise ScriptError(f"Variable not found: {name}", line)

    def tick(self, line=None):
        self.ops += 1
        if self.ops > self.MAX_OPERATIONS:
            raise ScriptError("Script exceeded the operation limit: too many operations", line)

    def exec(self, st):
        kind = st[0]
        self.tick()
        if kind == "expr":
            v = self.eval(st[1])
            return None if st[2] else v
        if kind == "let":
            self.scopes[-1][st[1]] = _copy_value(self.eval(st[2]))
            return None
        if kind == "assign":
            _, target, op, rhs, line = st
            val = self.eval(rhs)
            if target[0] == "var":
                sc = self.lookup(target[1], line)
                if self.vector_mode and id(sc) in self._outer_scopes:
                    raise _NotVectorisable()  # a captured variable is shared and mutated PER PIXEL in Rhai (counters, sums)
                sc[target[1]] = _copy_value(val if op == "=" else _binary(op[:-1], sc[target[1]], val, line))
            else:
                arr, idx = self.eval(target[1]), self.eval(target[2])
                if self.vector_mode and isinstance(arr, list) and id(arr) in self._outer_lists:
                    raise _NotVectorisable()  # e.g. hist[bin] += 1 on a captured array
                if not isinstance(arr, list) or not _is_int(idx) or _is_vec(idx):
                    if self.vector_mode:
                        raise _NotVectorisable()
                    raise ScriptError("Indexing assignment needs an array and an integer index", line)
                i = self.index_of(arr, idx, line)
                arr[i] = val if op == "=" else _binary(op[:-1], arr[i], val, line)
            return None
        if kind == "fn":
            self.functions[st[1]] = (st[2], st[3])
            return None
        if kind == "while":
            while True:
                c = self.eval(st[1])
                if _is_vec(c):
                    raise _NotVectorisable()
                if not self.truth(c, st[3]):
                    break
                try:
                    self.exec_block(st[2][1])
                except _Break:
                    break
                except _Continue:
                    continue
            return None
        if kind == "for":
            _, var, it, body, line = st
            seq = self.eval(it)
            if isinstance(seq, str):
                seq = list(seq)
            if _is_vec(seq) or not isinstance(seq, (range, list)):
                if self.vector_mode:
                    raise _NotVectorisable()
                raise ScriptError("For loop expects a range or an array", line)
            for v in list(seq):
                self.scopes.append({var: v})
                try:
                    self.exec_block(body[1], new_scope=False)
                except _Break:
                    break
                except _Continue:
                    continue
                finally:
                    self.scopes.pop()
                self.tick(line)
            return None
        if kind == "break":
            raise _Break()
        if kind == "continue":
            raise _Continue()
        if kind == "return":
            raise _Return(self.eval(st[1]))
        raise ScriptError(f"internal error: statement {kind}")

    def truth(self, v, line=None) -> bool:
        if isinstance(v, (bool, np.bool_)):
            return bool(v)
        raise ScriptError(f"Data type incorrect: bool expected, found {_type_name(v)}", line)

    def index_of(self, arr, idx, line=None) -> int:
        n = len(arr)
        i = idx + n if idx < 0 else idx
        if not 0 <= i < n:
            raise ScriptError(f"Array index {idx} out of bounds: only {n} elements in the array", line)
        return i

    # ---------------------------------------------------------------- expressions
    def eval(self, e):
        kind = e[0]
        if kind == "lit":
            return e[1]
        if kind == "var":
            return self.lookup(e[1], e[2])[e[1]]
        if kind == "unit":
            return None
        if kind == "bin":
            _, op, l, r, line = e
            if op in ("&&", "||"):
                a = self.eval(l)
                if _is_vec(a):
                    b = self.eval(r)
                    return np.logical_and(a, b) if op == "&&" else np.logical_or(a, b)
                a = self.truth(a, line)
                if (op == "&&") != a:
                    return a
                b = self.eval(r)
                return b if _is_vec(b) else self.truth(b, line)
            return _binary(op, self.eval(l), self.eval(r), line)
        if kind == "un":
            v = self.eval(e[2])
            if e[1] == "!":
                return np.logical_not(v) if _is_vec(v) else not self.truth(v, e[3])
            if not _is_num(v):
                raise ScriptError(f"Function not found: - ({_type_name(v)})", e[3])
            return -v
        if kind == "call":
            return self.call(e[1], [self.eval(a) for a in e[2]], e[3])
        if kind == "if":
            return self.eval_if(e)
        if kind == "block":
            return self.exec_block(e[1])
        if kind == "arr":
            return [self.eval(x) for x in e[1]]
        if kind == "index":
            obj, idx = self.eval(e[1]), self.eval(e[2])
            if _is_vec(idx):
                raise _NotVectorisable()
            if isinstance(obj, (list, str)) and _is_int(idx):
                return obj[self.index_of(obj, idx, e[3])]
            raise ScriptError(f"Indexing not supported on {_type_name(obj)}", e[3])
        if kind == "range":
            lo, hi = self.eval(e[1]), self.eval(e[2])
            if _is_vec(lo) or _is_vec(hi):
                raise _NotVectorisable()
            if not (_is_int(lo) and _is_int(hi)):
                raise ScriptError("Range bounds must be integers")
            return range(lo, hi + 1 if e[3] else hi)
        if kind == "tpl":
            parts = [self.eval(p) for p in e[1]]
            if any(_is_vec(p) for p in parts):
                raise _NotVectorisable()
            return "".join(to_text(p) for p in parts)
        if kind == "closure":
            return Closure(e[1], e[2], list(self.scopes))
        if kind == "method":
            return self.method(e)
        raise ScriptError(f"internal error: expression {kind}")

    def eval_if(self, e):
        _, cond, then, other, line = e
        c = self.eval(cond)
        if not _is_vec(c):
            if self.truth(c, line):
                return self.exec_block(then[1])
            if other is None:
                return None
            return self.eval_if(other) if other[0] == "if" else self.exec_block(other[1])
        # data-dependent branch over the whole image: run both arms on a copy of the variables, select per pixel
        if c.dtype != bool:
            raise ScriptError(f"Data type incorrect: bool expected, found {_type_name(c)}", line)
        if c.all():
            return self.exec_block(then[1])
        # The scope dicts are shared with closures that captured them, so they are rewound in place.
        saved = [{k: _copy_value(v) for k, v in sc.items()} for sc in self.scopes]
        try:
            v_then = self.exec_block(then[1])
            after_then = [dict(sc) for sc in self.scopes]
            for sc, old in zip(self.scopes, saved):
                sc.clear()
                sc.update(old)
            v_else = None if other is None else (self.eval_if(other) if other[0] == "if" else self.exec_block(other[1]))
        except (_Break, _Continue, _Return):
            raise _NotVectorisable()
        for sc_t, sc_e in zip(after_then, self.scopes):
            for k in sc_e:
                sc_e[k] = self.select(c, sc_t[k], sc_e[k])
        return self.select(c, v_then, v_else)

    def select(self, c, a, b):
        if a is b:
            return a
        if isinstance(a, list) and isinstance(b, list) and len(a) == len(b):
            return [self.select(c, x, y) for x, y in zip(a, b)]
        ta, tb = _type_name(a), _type_name(b)
        if ta == tb and ta in ("i64", "f64", "bool"):
            return np.where(c, a, b)
        if ta == tb and ta == "string" and a == b:
            return a
        raise _NotVectorisable()  # arms of different kinds (e.g. an array on one side, unit on the other)

    def method(self, e):
        _, obj_e, name, args_e, line = e
        obj = self.eval(obj_e)
        if args_e is None:
            raise ScriptError(f"Property {name} not found on {_type_name(obj)}", line)
        args = [self.eval(a) for a in args_e]
        if isinstance(obj, list):
            if self.vector_mode and name in ("push", "pop", "clear") and id(obj) in self._outer_lists:
                raise _NotVectorisable()  # a captured array grows / shrinks once per pixel in Rhai
            if name == "len":
                return len(obj)
            if name == "push":
                obj.append(args[0])
                return None
            if name == "pop":
                return obj.pop() if obj else None
            if name == "contains":
                return any(not _is_vec(x) and _binary("==", x, args[0]) for x in obj)
            if name == "clear":
                obj.clear()
                return None
            if name == "is_empty":
                return not obj
        if isinstance(obj, str):
            if name == "len":
                return len(obj)
            if name == "contains":
                return to_text(args[0]) in obj
            if name == "to_upper":
                return obj.upper()
            if name == "to_lower":
                return obj.lower()
        if isinstance(obj, Closure) and name == "call":
            return self.call_closure(obj, args, line)
        if name == "to_string":
            if _is_vec(obj):
                raise _NotVectorisable()
            return to_text(obj)
        if name == "type_of":
            return _type_name(obj)
        # obj.f(args) is f(obj, args) for registered and script functions
        return self.call(name, [obj] + args, line)

    def call_closure(self, fn: Closure, args, line=None):
        if len(args) != len(fn.params):
            raise ScriptError(f"Closure takes {len(fn.params)} arguments, {len(args)} given", line)
        saved, self.scopes = self.scopes, fn.scopes + [dict(zip(fn.params, args))]
        try:
            if fn.body[0] == "block":
                return self.exec_block(fn.body[1], new_scope=False)
            return self.eval(fn.body)
        except _Return as r:
            return r.value
        finally:
            self.scopes = saved

    def call(self, name, args, line=None):
        self.tick(line)
        if name in self.functions:
            params, body = self.functions[name]
            if len(params) == len(args):
                saved, self.scopes = self.scopes, [dict(zip(params, (_copy_value(a) for a in args)))]  # fn bodies see no outer variables
                try:
                    return self.exec_block(body[1], new_scope=False)
                except _Return as r:
                    return r.value
                finally:
                    self.scopes = saved
        for sc in reversed(self.scopes):  # a closure held in a variable
            if name in sc and isinstance(sc[name], Closure):
                return self.call_closure(sc[name], args, line)
        if name in _MATH:
            try:
                return _MATH[name](*args)
            except TypeError:
                raise ScriptError(f"Function not found: {name} ({', '.join(_type_name(a) for a in args)})", line)
        host = getattr(self, "api_" + name, None)
        if host is not None:
            try:
                return host(*args)
            except TypeError as err:
                if "positional argument" in str(err):
                    raise ScriptError(f"Function not found: {name} ({', '.join(_type_name(a) for a in args)})", line)
                raise
        if name in self.effects or name in self._script._SELECTION_API:
            if self.vector_mode:
                raise ScriptError(f"{name}() cannot be called from inside a per-pixel closure", line)
            if any(isinstance(a, (list, Closure)) or a is None for a in args):
                raise ScriptError(f"Function not found: {name} ({', '.join(_type_name(a) for a in args)})", line)
            return self.effect(name, args, line)
        raise ScriptError(f"Function not found: {name} ({', '.join(_type_name(a) for a in args)})", line)

    def effect(self, name, args, line):
        img = self.device_image()
        if name in self._script._SELECTION_API:
            img, self.mask = self._script._selection_call(self.eng, name, args, img, self.mask)
        else:
            h0, w0 = self.height, self.width
            try:
                img = self.effects[name](img, self.mask, *args)
            except TypeError as err:
                if "positional argument" in str(err):
                    raise ScriptError(f"Function not found: {name} ({', '.join(_type_name(a) for a in args)})", line)
                raise
            if name in self._script.CANVAS_OPS and not (name == "resize_image" and tuple(img.shape[:2]) == (h0, w0)):
                self.canvas_ops.append((name, tuple(args)))  # scripting.rs:687-813 (an unchanged-size resize logs nothing)
            if self.mask is not None and tuple(img.shape[:2]) != (h0, w0):
                self.mask = None  # the reference keeps a stale w*h mask after a resize / quarter turn; scripts re-select
        self.set_image(img)
        return None

    # ---------------------------------------------------------------- registered host API (scripting.rs:323-616, 1171-1481)
    def api_width(self):
        return self.width

    def api_height(self):
        return self.height

    def api_has_selection(self):
        return self.mask is not None

    def api_is_selected(self, x, y):
        m = self.host_mask()
        if _is_vec(x) or _is_vec(y):
            x, y = np.broadcast_arrays(np.asarray(x), np.asarray(y))
            inside = (x >= 0) & (y >= 0) & (x < self.width) & (y < self.height)
            if m is None:
                return inside
            return inside & (m[np.clip(y, 0, self.height - 1), np.clip(x, 0, self.width - 1)] > 0)
        if x < 0 or y < 0 or x >= self.width or y >= self.height:
            return False
        return True if m is None else bool(m[y, x] > 0)

    def _inside(self, x, y):
        if _is_vec(x) or _is_vec(y):
            raise _NotVectorisable()
        return 0 <= x < self.width and 0 <= y < self.height

    def api_get_pixel(self, x, y):
        if self.vector_mode:
            raise _NotVectorisable()  # reads of the image the closure is rewriting depend on the visiting order
        return [int(v) for v in self.host_pixels()[y, x]] if self._inside(x, y) else [0, 0, 0, 0]

    def api_set_pixel(self, x, y, r, g, b, a):
        if self.vector_mode:
            raise _NotVectorisable()
        if self._inside(x, y):
            self.touch_host()
            self.host_pixels()[y, x] = [min(max(int(v), 0), 255) for v in (r, g, b, a)]

    def _get_channel(self, ch, x, y):
        if self.vector_mode:
            raise _NotVectorisable()
        return int(self.host_pixels()[y, x, ch]) if self._inside(x, y) else 0

    def _set_channel(self, ch, x, y, v):
        if self.vector_mode:
            raise _NotVectorisable()
        if self._inside(x, y):
            self.touch_host()
            self.host_pixels()[y, x, ch] = min(max(int(v), 0), 255)

    api_get_r = lambda self, x, y: self._get_channel(0, x, y)
    api_get_g = lambda self, x, y: self._get_channel(1, x, y)
    api_get_b = lambda self, x, y: self._get_channel(2, x, y)
    api_get_a = lambda self, x, y: self._get_channel(3, x, y)
    api_set_r = lambda self, x, y, v: self._set_channel(0, x, y, v)
    api_set_g = lambda self, x, y, v: self._set_channel(1, x, y, v)
    api_set_b = lambda self, x, y, v: self._set_channel(2, x, y, v)
    api_set_a = lambda self, x, y, v: self._set_channel(3, x, y, v)

    def _say(self, msg):
        if self.vector_mode:
            raise _NotVectorisable()
        self.console.append(to_text(msg))

    api_print = api_print_line = api_debug = _say

    def api_sleep(self, ms):
        return None  # preview + pause for the GUI (scripting.rs:1189); nothing to show in batch mode

    def api_progress(self, frac):
        return None

    def _xorshift(self) -> int:
        if self.vector_mode:
            raise _NotVectorisable()  # the sequence depends on the visiting order
        s = self.rng_state
        s ^= (s << 13) & 0xFFFFFFFFFFFFFFFF
        s ^= s >> 7
        s ^= (s << 17) & 0xFFFFFFFFFFFFFFFF
        self.rng_state = s
        return s

    def api_rand_int(self, lo, hi):  # scripting.rs:1216
        return lo if lo >= hi else lo + self._xorshift() % max(hi - lo, 1)

    def api_rand_float(self, lo=0.0, hi=1.0):  # scripting.rs:1232 / :1247
        if lo >= hi:
            return float(lo)
        return float(lo) + (self._xorshift() / float(0xFFFFFFFFFFFFFFFF)) * (float(hi) - float(lo))

    def api_rgb_to_hsl(self, r, g, b):
        if _is_vec(r) or _is_vec(g) or _is_vec(b):
            raise _NotVectorisable()
        return _rgb_to_hsl(r, g, b)

    def api_hsl_to_rgb(self, h, s, l):
        if _is_vec(h) or _is_vec(s) or _is_vec(l):
            raise _NotVectorisable()
        return _hsl_to_rgb(h, s, l)

    def api_range(self, lo, hi, step=1):
        if step == 0:
            raise ScriptError("range: step is zero")
        return range(lo, hi, step)

    # -- bulk iteration (scripting.rs:437-616)
    def api_for_each_pixel(self, fn):
        self._bulk(fn, 0, 0, self.width, self.height, True)

    def api_for_region(self, rx, ry, rw, rh, fn):
        wrap = lambda v: v & 0xFFFFFFFF  # `(rx + rw) as u32`
        x0, y0 = max(rx, 0), max(ry, 0)
        x1, y1 = min(wrap(rx + rw), self.width), min(wrap(ry + rh), self.height)
        if x1 > x0 and y1 > y0:
            self._bulk(fn, x0, y0, x1, y1, True)

    def api_map_channels(self, fn):
        self._bulk(fn, 0, 0, self.width, self.height, False)

    def _bulk(self, fn, x0, y0, x1, y1, with_xy):
        if not isinstance(fn, Closure):
            raise ScriptError("expected a closure: |x, y, r, g, b, a| { ... }")
        if self.vector_mode:
            raise ScriptError("bulk pixel iteration cannot be nested inside a per-pixel closure")
        src = self.host_pixels()
        region = src[y0:y1, x0:x1].astype(np.int64)
        out = None
        saved_ops, saved_console, saved_rng = self.ops, len(self.console), self.rng_state
        # Only a PURE closure may be evaluated once over whole-image arrays: writes to anything outside its own call
        # frame (captured variables, arrays reachable from them) abort the attempt before they happen, so the
        # per-pixel evaluation below starts from untouched state.
        outer = list(fn.scopes)
        self._outer_scopes = frozenset(id(sc) for sc in outer)
        lists, stack = set(), [v for sc in outer for v in sc.values()]
        while stack:
            v = stack.pop()
            if isinstance(v, list) and id(v) not in lists:
                lists.add(id(v))
                stack.extend(v)
            elif isinstance(v, Closure):
                stack.extend(x for sc in v.scopes for x in sc.values() if isinstance(x, list))
        self._outer_lists = frozenset(lists)
        try:
            self.vector_mode = True
            ys, xs = np.mgrid[y0:y1, x0:x1].astype(np.int64)
            chans = [np.ascontiguousarray(region[..., c]) for c in range(4)]
            res = self.call_closure(fn, ([xs, ys] if with_xy else []) + chans)
            out = self._store(region, res)
        except (_NotVectorisable, ScriptError, ValueError, TypeError, IndexError):
            out = None  # evaluate per pixel instead; a genuine script error is raised again there
            del self.console[saved_console:]
            self.ops, self.rng_state = saved_ops, saved_rng
        finally:
            self.vector_mode = False
            self._outer_scopes = self._outer_lists = frozenset()
        self.bulk_evaluations["whole_image" if out is not None else "per_pixel"] += 1
        if out is None:
            out = self._bulk_scalar(fn, region, x0, y0, with_xy)
        self.touch_host()
        self.host_pixels()[y0:y1, x0:x1] = out
        if not isinstance(self.img, np.ndarray):
            self._host_dirty = True

    @staticmethod
    def _store(region, res):
        """`arr[i].as_int().unwrap_or(old).clamp(0, 255)`: integer results are written, anything else keeps the channel."""
        out = region.astype(np.uint8)
        if not (isinstance(res, list) and len(res) >= 4):
            if _is_vec(res):
                raise _NotVectorisable()
            return out
        for c in range(4):
            v = res[c]
            if _is_vec(v) and v.dtype == bool:
                continue
            if _is_int(v):
                out[..., c] = np.clip(np.broadcast_to(v, region.shape[:2]), 0, 255).astype(np.uint8)
            elif isinstance(v, list):
                raise _NotVectorisable()
        return out

    def _bulk_scalar(self, fn, region, x0, y0, with_xy):
        out = region.astype(np.uint8)
        h, w = region.shape[:2]
        # the reference's closures read a snapshot (get_pixel inside sees the image as it was before the call)
        for yy in range(h):
            for xx in range(w):
                px = [int(v) for v in region[yy, xx]]
                res = self.call_closure(fn, ([x0 + xx, y0 + yy] if with_xy else []) + px)
                if isinstance(res, list) and len(res) >= 4:
                    out[yy, xx] = [min(max(res[c], 0), 255) if _is_int(res[c]) else px[c] for c in range(4)]
        return out


def run_script(eng, source: str, pixels, mask=None, exact: bool = False):
    """(image, console lines) - the Ok arm of scripting::execute_script_sync (scripting.rs:1733)."""
    it = Interpreter(eng, pixels, mask, exact)
    img = it.run(source)
    return img, it.console

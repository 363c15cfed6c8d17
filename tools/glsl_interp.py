"""A small GLSL ES 1.00 interpreter -- just enough of the language to EXECUTE THE REFERENCE'S OWN SHADER
TEXT (the glslified sources embedded in /root/reference/docs/js/*.js.map) one fragment / vertex at a time.

Purpose: pin the CPU oracle.  The reference has no tests and no GL implementation exists in this image, so
instead of trusting a hand transcription alone, tools/make_glsl_golden.py runs the real shader source through
this interpreter and commits inputs + outputs under tests/golden/; tests/test_glsl_golden.py then demands that
oracle/tendrils_oracle.c reproduces them bit for bit.  The interpreter shares no code with the oracle.

Arithmetic model (spec/PARITY.md R1-R6): every operator is one binary32 operation (numpy float32), evaluated
in source order; min/max/step/mod/fract/mix by their GLSL defining formulas; sin/cos by TSIN-1 (implemented
here a third time, in numpy); texture2D is supplied by the caller.

Supported: global uniform/varying/attribute/const declarations, function definitions with overloading,
float/vecN/matN/bool/int locals, if/else, for, return, assignment operators incl. swizzled l-values,
ternary, constructors, swizzles, matN*vecN, the built-ins the path uses.
"""
from __future__ import annotations

import re

import numpy as np

f32 = np.float32
np.seterr(all="ignore")

TYPES = {"void", "float", "int", "bool", "vec2", "vec3", "vec4", "mat2", "mat3", "mat4", "sampler2D"}
QUALS = {"uniform", "varying", "attribute", "const", "highp", "mediump", "lowp", "in", "out", "inout"}
VEC_N = {"vec2": 2, "vec3": 3, "vec4": 4}
MAT_N = {"mat2": 2, "mat3": 3, "mat4": 4}
SWZ = {c: i for s in ("xyzw", "rgba", "stpq") for i, c in enumerate(s)}

TOKEN = re.compile(r"\s*(?:(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)|(\d+)|([A-Za-z_]\w*)|"
                   r"(\+\+|--|\+=|-=|\*=|/=|<=|>=|==|!=|&&|\|\||[-+*/<>=!?:;,.(){}\[\]]))")


def tokenize(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    lines, keep = [], [True]                              # minimal preprocessor: #if <int> / #else / #endif
    for l in src.split("\n"):
        st = l.strip()
        if st.startswith("#if"):
            keep.append(keep[-1] and bool(int(st.split()[1])))
        elif st.startswith("#else"):
            keep[-1] = (not keep[-1]) and keep[-2]
        elif st.startswith("#endif"):
            keep.pop()
        elif st.startswith("#"):
            pass
        elif keep[-1]:
            lines.append(l)
    src = "\n".join(lines)
    out, pos = [], 0
    while True:
        m = TOKEN.match(src, pos)
        if not m:
            if src[pos:].strip():
                raise SyntaxError("cannot tokenize: %r" % src[pos:pos + 40])
            break
        pos = m.end()
        fl, it, idn, op = m.groups()
        if fl is not None:
            out.append(("float", f32(fl)))
        elif it is not None:
            out.append(("int", int(it)))
        elif idn is not None:
            out.append(("id", idn))
        else:
            out.append(("op", op))
    out.append(("eof", None))
    return out


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def accept(self, kind, val=None):
        k, v = self.peek()
        if k == kind and (val is None or v == val):
            self.i += 1
            return True
        return False

    def expect(self, kind, val=None):
        k, v = self.next()
        if k != kind or (val is not None and v != val):
            raise SyntaxError(f"expected {val or kind}, got {v!r} at token {self.i}")
        return v

    # ---- declarations -------------------------------------------------------------------------
    def unit(self):
        items = []
        while self.peek()[0] != "eof":
            if self.peek() == ("id", "precision"):
                while not self.accept("op", ";"):
                    self.next()
                continue
            items.append(self.external())
        return items

    def quals_type(self):
        quals = []
        while self.peek()[0] == "id" and self.peek()[1] in QUALS:
            quals.append(self.next()[1])
        ty = self.expect("id")
        if ty not in TYPES:
            raise SyntaxError("unknown type " + ty)
        return quals, ty

    def external(self):
        quals, ty = self.quals_type()
        name = self.expect("id")
        if self.accept("op", "("):
            params = []
            if not self.accept("op", ")"):
                while True:
                    if self.peek() == ("id", "void") and self.peek(1) == ("op", ")"):
                        self.next()
                        break
                    _, pty = self.quals_type()
                    params.append((pty, self.expect("id")))
                    if not self.accept("op", ","):
                        break
                self.expect("op", ")")
            if self.accept("op", ";"):
                return ("proto",)
            return ("func", ty, name, params, self.compound())
        decls = [self.declarator(name)]
        while self.accept("op", ","):
            decls.append(self.declarator(self.expect("id")))
        self.expect("op", ";")
        return ("global", quals, ty, decls)

    def declarator(self, name):
        init = self.assign() if self.accept("op", "=") else None
        return (name, init)

    # ---- statements ---------------------------------------------------------------------------
    def compound(self):
        self.expect("op", "{")
        body = []
        while not self.accept("op", "}"):
            body.append(self.statement())
        return ("block", body)

    def statement(self):
        k, v = self.peek()
        if (k, v) == ("op", "{"):
            return self.compound()
        if (k, v) == ("op", ";"):
            self.next()
            return ("block", [])
        if k == "id" and v == "if":
            self.next()
            self.expect("op", "(")
            c = self.expr()
            self.expect("op", ")")
            a = self.statement()
            b = self.statement() if self.accept("id", "else") else None
            return ("if", c, a, b)
        if k == "id" and v == "for":
            self.next()
            self.expect("op", "(")
            init = self.statement()
            cond = self.expr()
            self.expect("op", ";")
            step = self.expr()
            self.expect("op", ")")
            return ("for", init, cond, step, self.statement())
        if k == "id" and v == "return":
            self.next()
            e = None if self.peek() == ("op", ";") else self.expr()
            self.expect("op", ";")
            return ("return", e)
        if k == "id" and (v in TYPES or v in QUALS):
            _, ty = self.quals_type()
            decls = [self.declarator(self.expect("id"))]
            while self.accept("op", ","):
                decls.append(self.declarator(self.expect("id")))
            self.expect("op", ";")
            return ("decl", ty, decls)
        e = self.expr()
        self.expect("op", ";")
        return ("expr", e)

    # ---- expressions --------------------------------------------------------------------------
    def expr(self):
        return self.assign()

    def assign(self):
        lhs = self.ternary()
        k, v = self.peek()
        if k == "op" and v in ("=", "+=", "-=", "*=", "/="):
            self.next()
            return ("assign", v, lhs, self.assign())
        return lhs

    def ternary(self):
        c = self.binary(0)
        if self.accept("op", "?"):
            a = self.assign()
            self.expect("op", ":")
            b = self.assign()
            return ("ternary", c, a, b)
        return c

    LEVELS = [("||",), ("&&",), ("==", "!="), ("<", ">", "<=", ">="), ("+", "-"), ("*", "/")]

    def binary(self, lvl):
        if lvl == len(self.LEVELS):
            return self.unary()
        lhs = self.binary(lvl + 1)
        while self.peek()[0] == "op" and self.peek()[1] in self.LEVELS[lvl]:
            op = self.next()[1]
            lhs = ("bin", op, lhs, self.binary(lvl + 1))
        return lhs

    def unary(self):
        if self.accept("op", "-"):
            return ("neg", self.unary())
        if self.accept("op", "+"):
            return self.unary()
        if self.accept("op", "!"):
            return ("not", self.unary())
        return self.postfix()

    def postfix(self):
        k, v = self.next()
        if k in ("float", "int"):
            e = ("lit", v)
        elif (k, v) == ("op", "("):
            e = self.expr()
            self.expect("op", ")")
        elif k == "id":
            if self.accept("op", "("):
                args = []
                if not self.accept("op", ")"):
                    while True:
                        args.append(self.assign())
                        if not self.accept("op", ","):
                            break
                    self.expect("op", ")")
                e = ("call", v, args)
            else:
                e = ("var", v)
        else:
            raise SyntaxError(f"unexpected {v!r}")
        while True:
            if self.accept("op", "."):
                e = ("field", e, self.expect("id"))
            elif self.accept("op", "["):
                idx = self.expr()
                self.expect("op", "]")
                e = ("index", e, idx)
            elif self.accept("op", "++"):
                e = ("assign", "+=", e, ("lit", f32(1)))
            else:
                return e


class Return(Exception):
    def __init__(self, v):
        self.v = v


def tyof(v):
    if isinstance(v, (bool, np.bool_)):
        return "bool"
    if isinstance(v, int):
        return "int"
    if isinstance(v, np.ndarray):
        if v.ndim == 2:
            return {2: "mat2", 3: "mat3", 4: "mat4"}[v.shape[0]]
        return {2: "vec2", 3: "vec3", 4: "vec4"}[v.shape[0]]
    if callable(v):
        return "sampler2D"
    return "float"


def tsin1(x):
    """TSIN-1 (spec/PARITY.md R6) in numpy float32; returns (sin, cos)."""
    x = f32(x)
    if not (abs(x) <= f32(100000.0)):
        return f32(np.nan), f32(np.nan)
    kf = x * f32(0.636619772)
    kf = (kf + f32(12582912.0)) - f32(12582912.0)
    r = x - kf * f32(1.5703125)
    r = r - kf * f32(4.837512969970703125e-4)
    r = r - kf * f32(7.54978995489188216e-8)
    q = int(kf) & 3
    z = r * r
    s = ((((f32(-1.9515295891e-4) * z + f32(8.3321608736e-3)) * z - f32(1.6666654611e-1)) * z) * r) + r
    c = ((((f32(2.443315711809948e-5) * z - f32(1.388731625493765e-3)) * z + f32(4.166664568298827e-2)) * z) * z
         - f32(0.5) * z) + f32(1.0)
    return [(s, c), (c, -s), (-s, -c), (-c, s)][q]


def _cw(fn):
    """component-wise over scalars / vectors with scalar broadcast"""
    def g(*args):
        n = max((a.shape[0] for a in args if isinstance(a, np.ndarray)), default=0)
        if n == 0:
            return f32(fn(*[f32(a) for a in args]))
        cols = [a if isinstance(a, np.ndarray) else np.full(n, a, f32) for a in args]
        return np.array([fn(*[c[i] for c in cols]) for i in range(n)], f32)
    return g


def _dot(a, b):
    if not isinstance(a, np.ndarray):
        return f32(a * b)
    acc = a[0] * b[0]
    for i in range(1, a.shape[0]):
        acc = f32(acc + a[i] * b[i])
    return f32(acc)


BUILTINS = {
    "floor": _cw(lambda x: np.floor(x)),
    "fract": _cw(lambda x: x - np.floor(x)),
    "abs": _cw(lambda x: np.abs(x)),
    "sqrt": _cw(lambda x: np.sqrt(x)),
    "sin": _cw(lambda x: tsin1(x)[0]),
    "cos": _cw(lambda x: tsin1(x)[1]),
    "mod": _cw(lambda x, y: x - y * np.floor(x / y)),
    "min": _cw(lambda x, y: y if y < x else x),
    "max": _cw(lambda x, y: y if x < y else x),
    "step": _cw(lambda e, x: f32(0.0) if x < e else f32(1.0)),
    "mix": _cw(lambda x, y, a: x * (f32(1.0) - a) + y * a),
    "dot": _dot,
    "length": lambda v: f32(np.sqrt(_dot(v, v))),
    # GLSL ES 1.00 8.3: 1.0 if x > 0, 0.0 if x = 0, -1.0 if x < 0 (NaN falls through to 0, spec/PARITY.md FL2)
    "sign": _cw(lambda x: f32(1.0) if x > 0 else (f32(-1.0) if x < 0 else f32(0.0))),
    # v / length(v), component-wise (spec/PARITY.md FL2); the zero vector gives 0/0 = NaN
    "normalize": lambda v: (v / f32(np.sqrt(_dot(v, v)))).astype(f32),
}


class Shader:
    def __init__(self, source):
        self.ast = Parser(tokenize(source)).unit()
        self.funcs, self.globals_ast = {}, []
        for item in self.ast:
            if item[0] == "func":
                _, ty, name, params, body = item
                self.funcs.setdefault(name, []).append((tuple(p[0] for p in params), [p[1] for p in params], body, ty))
            elif item[0] == "global":
                self.globals_ast.append(item)

    # ---- running ------------------------------------------------------------------------------
    def run(self, inputs):
        """inputs: uniforms / attributes / gl_FragCoord by name.  Returns the global scope after main()."""
        self.g = {}
        for _, quals, ty, decls in self.globals_ast:
            for name, init in decls:
                if init is not None:
                    self.g[name] = self.eval(init, [self.g])
                elif name in inputs:
                    self.g[name] = self.coerce(ty, inputs[name])
                else:
                    self.g[name] = self.zero(ty)
        for k, v in inputs.items():
            if k.startswith("gl_"):
                self.g[k] = np.asarray(v, f32)
        self.call("main", [])
        return self.g

    @staticmethod
    def zero(ty):
        if ty in VEC_N:
            return np.zeros(VEC_N[ty], f32)
        if ty in MAT_N:
            return np.zeros((MAT_N[ty], MAT_N[ty]), f32)
        return f32(0)

    @staticmethod
    def coerce(ty, v):
        if ty == "sampler2D":
            return v
        if ty in VEC_N:
            return np.asarray(v, f32).reshape(VEC_N[ty]).copy()
        if ty in MAT_N:
            n = MAT_N[ty]
            return np.asarray(v, f32).reshape(n, n).copy()      # column-major: m[col][row]
        if ty == "bool":
            return bool(v)
        if ty == "int":
            return int(v)
        return f32(v)

    def call(self, name, args):
        if name in self.funcs:
            sig = tuple(tyof(a) for a in args)
            for psig, pnames, body, _ in self.funcs[name]:
                if psig == sig:
                    scope = {n: (a.copy() if isinstance(a, np.ndarray) else a) for n, a in zip(pnames, args)}
                    try:
                        self.exec(body, [self.g, scope])
                    except Return as r:
                        return r.v
                    return None
            raise TypeError(f"no overload {name}{sig}")
        if name in VEC_N or name in MAT_N or name in ("float", "int", "bool"):
            return self.construct(name, args)
        if name == "texture2D":
            return np.asarray(args[0](args[1][0], args[1][1]), f32)
        if name in BUILTINS:
            return BUILTINS[name](*args)
        raise NameError(name)

    @staticmethod
    def construct(ty, args):
        if ty == "float":
            a = args[0]
            return f32(a[0] if isinstance(a, np.ndarray) else a)
        if ty in VEC_N:
            n = VEC_N[ty]
            flat = []
            for a in args:
                flat.extend(list(a.reshape(-1)) if isinstance(a, np.ndarray) else [f32(a)])
            if len(flat) == 1:
                flat = flat * n
            return np.array(flat[:n], f32)
        raise TypeError(ty)

    # ---- statements ---------------------------------------------------------------------------
    def exec(self, node, env):
        kind = node[0]
        if kind == "block":
            env = env + [{}]
            for s in node[1]:
                self.exec(s, env)
        elif kind == "decl":
            _, ty, decls = node
            for name, init in decls:
                v = self.eval(init, env) if init is not None else self.zero(ty)
                env[-1][name] = self.coerce(ty, v) if ty != "sampler2D" else v
        elif kind == "expr":
            self.eval(node[1], env)
        elif kind == "if":
            if self.eval(node[1], env):
                self.exec(node[2], env)
            elif node[3] is not None:
                self.exec(node[3], env)
        elif kind == "for":
            env = env + [{}]
            self.exec(node[1], env)
            while self.eval(node[2], env):
                self.exec(node[4], env)
                self.eval(node[3], env)
        elif kind == "return":
            v = self.eval(node[1], env) if node[1] is not None else None
            raise Return(v.copy() if isinstance(v, np.ndarray) else v)
        else:
            raise NotImplementedError(kind)

    # ---- expressions --------------------------------------------------------------------------
    def lookup(self, name, env):
        for scope in reversed(env):
            if name in scope:
                return scope
        raise NameError(name)

    def eval(self, node, env):
        kind = node[0]
        if kind == "lit":
            return node[1]
        if kind == "var":
            return self.lookup(node[1], env)[node[1]]
        if kind == "neg":
            return -self.eval(node[1], env)
        if kind == "not":
            return not self.eval(node[1], env)
        if kind == "bin":
            return self.binop(node[1], self.eval(node[2], env), self.eval(node[3], env))
        if kind == "ternary":
            return self.eval(node[2], env) if self.eval(node[1], env) else self.eval(node[3], env)
        if kind == "call":
            return self.call(node[1], [self.eval(a, env) for a in node[2]])
        if kind == "field":
            v = self.eval(node[1], env)
            idx = [SWZ[c] for c in node[2]]
            return f32(v[idx[0]]) if len(idx) == 1 else v[idx].copy()
        if kind == "index":
            v = self.eval(node[1], env)
            i = int(self.eval(node[2], env))
            return v[i].copy() if v.ndim == 2 else f32(v[i])
        if kind == "assign":
            return self.assign(node[1], node[2], self.eval(node[3], env), env)
        raise NotImplementedError(kind)

    @staticmethod
    def binop(op, a, b):
        if op in ("&&", "||"):
            return (a and b) if op == "&&" else (a or b)
        if op in ("==", "!="):
            eq = bool(np.all(np.asarray(a) == np.asarray(b)))
            return eq if op == "==" else not eq
        if op in ("<", ">", "<=", ">="):
            return bool({"<": a < b, ">": a > b, "<=": a <= b, ">=": a >= b}[op])
        if isinstance(a, int) and isinstance(b, int):
            return {"+": a + b, "-": a - b, "*": a * b, "/": a // b}[op]
        am, bm = isinstance(a, np.ndarray) and a.ndim == 2, isinstance(b, np.ndarray) and b.ndim == 2
        if op == "*" and am and isinstance(b, np.ndarray) and b.ndim == 1:     # matN * vecN, columns weighted by v
            n = b.shape[0]
            out = np.zeros(n, f32)
            for r in range(n):
                acc = a[0][r] * b[0]
                for c in range(1, n):
                    acc = f32(acc + a[c][r] * b[c])
                out[r] = acc
            return out
        if am or bm:
            raise NotImplementedError("matrix op")
        a = a if isinstance(a, np.ndarray) else f32(a)
        b = b if isinstance(b, np.ndarray) else f32(b)
        r = {"+": lambda: a + b, "-": lambda: a - b, "*": lambda: a * b, "/": lambda: a / b}[op]()
        return r.astype(f32) if isinstance(r, np.ndarray) else f32(r)

    def assign(self, op, target, val, env):
        if op != "=":
            val = self.binop(op[0], self.eval(target, env), val)
        if target[0] == "var":
            scope = self.lookup(target[1], env) if any(target[1] in s for s in env) else self.g
            cur = scope.get(target[1])
            if isinstance(cur, np.ndarray):
                val = np.broadcast_to(np.asarray(val, f32), cur.shape).copy()
            scope[target[1]] = val.copy() if isinstance(val, np.ndarray) else val
        elif target[0] == "field":
            base = self.eval(target[1], env)         # ndarray: mutate in place through its owner
            owner = target[1]
            if owner[0] != "var":
                raise NotImplementedError("nested l-value")
            arr = self.lookup(owner[1], env)[owner[1]]
            idx = [SWZ[c] for c in target[2]]
            arr[idx] = val
            del base
        else:
            raise NotImplementedError("l-value " + target[0])
        return val


def nearest_sampler(tex):
    """texture2D for an [h, w, 4] float32 array: NEAREST, CLAMP_TO_EDGE (spec/PARITY.md T1, T2)."""
    h, w = tex.shape[:2]

    def texel(u, size):
        f = np.floor(f32(u) * f32(size))
        if not (f > 0):
            return 0
        if f > size - 1:
            return size - 1
        return int(f)

    return lambda u, v: tex[texel(v, h), texel(u, w)]

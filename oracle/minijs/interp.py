"""minijs — a small ECMAScript-subset interpreter, written to execute the UNMODIFIED
/root/reference/main.js in a container that has no JavaScript engine (no node / deno / bun / V8).

TEST INFRASTRUCTURE ONLY (see oracle/gen_golden.py).  It implements exactly the language surface
main.js uses — function declarations/expressions with closures and hoisting, rest parameters,
const/let/var, object and array literals, member access and calls with `this`, `new`, the arithmetic,
comparison, logical, bitwise (& ^ | << >>), conditional, assignment and postfix/prefix update
operators, if / for / while / return / break / continue — with JS number semantics: every number is
an IEEE-754 binary64 (a Python float), `&`/`^` go through ToInt32, `==` is loose equality, arrays grow
on out-of-range stores, `Array.prototype.sort` is stable.  `Math.*` maps to the host libm (V8 uses an
fdlibm port: atan2/asin/pow/tan may differ from it in the last ulp — stated in DESIGN.md).

Simplification that is safe for main.js: `let`/`const` are function-scoped (no closure in main.js
captures a per-iteration loop binding).

The AST is compiled to Python closures once; evaluation is then plain closure calls.
"""
from __future__ import annotations

import math
import re

# ---- values ------------------------------------------------------------------------------------------------


class _Undefined:
    def __repr__(self):
        return "undefined"

    def __bool__(self):
        return False


UNDEF = _Undefined()


class JSObject:
    __slots__ = ("props",)

    def __init__(self, props=None):
        self.props = props if props is not None else {}

    def js_get(self, k):
        return self.props.get(k, UNDEF)

    def js_set(self, k, v):
        self.props[k] = v


class JSArray:
    __slots__ = ("items", "props")

    def __init__(self, items=None):
        self.items = items if items is not None else []
        self.props = {}

    def js_get(self, k):
        if isinstance(k, (int, float)):
            i = int(k)
            if i == k and 0 <= i < len(self.items):
                return self.items[i]
            return UNDEF
        if k == "length":
            return float(len(self.items))
        m = _ARRAY_METHODS.get(k)
        if m is not None:
            return NativeFunction(lambda this, args, m=m: m(this, args), k)
        return self.props.get(k, UNDEF)

    def js_set(self, k, v):
        if isinstance(k, (int, float)) and int(k) == k and k >= 0:
            i = int(k)
            n = len(self.items)
            if i >= n:                                  # arrays grow on out-of-range stores
                self.items.extend([UNDEF] * (i - n + 1))
            self.items[i] = v
        elif k == "length":
            del self.items[int(v):]
        else:
            self.props[k] = v


class Uint8ClampedArray:
    """Typed array with the clamped store: NaN -> 0, clamp to [0,255], round half to even."""
    __slots__ = ("buf",)

    def __init__(self, n_or_bytes):
        self.buf = bytearray(n_or_bytes)

    def js_get(self, k):
        if isinstance(k, (int, float)):
            i = int(k)
            if i == k and 0 <= i < len(self.buf):
                return float(self.buf[i])
            return UNDEF
        if k == "length":
            return float(len(self.buf))
        return UNDEF

    def js_set(self, k, v):
        if isinstance(k, (int, float)) and int(k) == k and 0 <= k < len(self.buf):
            v = to_number(v)
            if not (v > 0):
                b = 0
            elif v >= 255:
                b = 255
            else:
                b = int(round(v))                       # Python rounds half to even, like the spec
            self.buf[int(k)] = b


class JSFunction:
    __slots__ = ("params", "rest", "body", "env", "name", "props", "hoist")

    def __init__(self, params, rest, body, env, name, hoist):
        self.params, self.rest, self.body, self.env, self.name = params, rest, body, env, name
        self.hoist = hoist
        self.props = {}

    def js_get(self, k):
        if k == "apply":
            return NativeFunction(lambda this, args: call(this, args[0] if args else UNDEF,
                                                          list(args[1].items) if len(args) > 1 else []), "apply")
        if k == "call":
            return NativeFunction(lambda this, args: call(this, args[0] if args else UNDEF, list(args[1:])), "call")
        return self.props.get(k, UNDEF)

    def js_set(self, k, v):
        self.props[k] = v


class NativeFunction:
    __slots__ = ("fn", "name", "construct")

    def __init__(self, fn, name="native", construct=None):
        self.fn, self.name, self.construct = fn, name, construct

    def js_get(self, k):
        if k == "apply":
            return NativeFunction(lambda this, args: self.fn(args[0] if args else UNDEF,
                                                             list(args[1].items) if len(args) > 1 else []), "apply")
        return UNDEF

    def js_set(self, k, v):
        pass


class _Return(Exception):
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = value


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class JSError(Exception):
    pass


# ---- conversions ----------------------------------------------------------------------------------------------

def truthy(v):
    if v is UNDEF or v is None or v is False:
        return False
    if isinstance(v, (int, float)):
        return not (v == 0 or v != v)
    if isinstance(v, str):
        return len(v) > 0
    return True


def to_number(v):
    if isinstance(v, bool):
        return 1.0 if v else 0.0
    if isinstance(v, (int, float)):
        return v
    if v is UNDEF:
        return math.nan
    if v is None:
        return 0.0
    if isinstance(v, str):
        try:
            return float(v) if v.strip() else 0.0
        except ValueError:
            return math.nan
    return math.nan


def to_int32(v):
    d = to_number(v)
    if d != d or d in (math.inf, -math.inf):
        return 0
    m = math.fmod(math.trunc(d), 4294967296.0)
    if m < 0:
        m += 4294967296.0
    u = int(m)
    return u - (1 << 32) if u >= (1 << 31) else u


def to_string(v):
    if isinstance(v, str):
        return v
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, float)):
        if v != v:
            return "NaN"
        if v in (math.inf, -math.inf):
            return "Infinity" if v > 0 else "-Infinity"
        if float(v).is_integer() and abs(v) < 1e21:
            return str(int(v))
        return repr(float(v))
    if v is UNDEF:
        return "undefined"
    if v is None:
        return "null"
    if isinstance(v, JSArray):
        return ",".join("" if x is UNDEF or x is None else to_string(x) for x in v.items)
    if isinstance(v, (JSFunction, NativeFunction)):
        return "function"
    return "[object Object]"


def js_typeof(v):
    if v is UNDEF:
        return "undefined"
    if isinstance(v, bool):
        return "boolean"
    if isinstance(v, (int, float)):
        return "number"
    if isinstance(v, str):
        return "string"
    if isinstance(v, (JSFunction, NativeFunction)):
        return "function"
    return "object"


def loose_eq(a, b):
    na, nb = isinstance(a, (int, float)), isinstance(b, (int, float))
    if na and nb:
        return a == b
    if (a is UNDEF or a is None) and (b is UNDEF or b is None):
        return True
    if (a is UNDEF or a is None) or (b is UNDEF or b is None):
        return False
    if isinstance(a, str) and isinstance(b, str):
        return a == b
    if na and isinstance(b, str):
        return a == to_number(b)
    if nb and isinstance(a, str):
        return to_number(a) == b
    return a is b


def strict_eq(a, b):
    na, nb = isinstance(a, (int, float)) and not isinstance(a, bool), isinstance(b, (int, float)) and not isinstance(b, bool)
    if na and nb:
        return a == b
    if type(a) is not type(b):
        return False
    if isinstance(a, (str, bool)):
        return a == b
    return a is b


def js_add(a, b):
    if isinstance(a, str) or isinstance(b, str):
        return to_string(a) + to_string(b)
    if isinstance(a, (JSObject, JSArray)) or isinstance(b, (JSObject, JSArray)):
        return to_string(a) + to_string(b)
    return to_number(a) + to_number(b)


def js_div(a, b):
    a, b = to_number(a), to_number(b)
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return math.nan
        neg = (a < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf


def get_member(obj, k):
    if isinstance(obj, str):
        if k == "length":
            return float(len(obj))
        raise JSError(f"string property {k!r} not supported")
    if obj is UNDEF or obj is None:
        raise JSError(f"cannot read property {k!r} of {obj!r}")
    return obj.js_get(k)


def call(fn, this, args):
    if isinstance(fn, JSFunction):
        env = {"this": this}
        for name in fn.hoist:                       # var/let/const/function names of the body
            env[name] = UNDEF
        ps = fn.params
        na = len(args)
        for i, p in enumerate(ps):
            env[p] = args[i] if i < na else UNDEF
        if fn.rest:
            env[fn.rest] = JSArray(list(args[len(ps):]))
        scope = Scope(env, fn.env)
        try:
            fn.body(scope)
        except _Return as r:
            return r.value
        return UNDEF
    if isinstance(fn, NativeFunction):
        return fn.fn(this, args)
    raise JSError(f"{fn!r} is not a function")


def construct(fn, args):
    if isinstance(fn, NativeFunction) and fn.construct is not None:
        return fn.construct(args)
    if isinstance(fn, JSFunction):
        obj = JSObject()
        r = call(fn, obj, args)
        return r if isinstance(r, (JSObject, JSArray)) else obj
    raise JSError("not a constructor")


class Scope:
    __slots__ = ("vars", "parent")

    def __init__(self, vars_, parent):
        self.vars, self.parent = vars_, parent

    def lookup(self, name):
        s = self
        while s is not None:
            v = s.vars
            if name in v:
                return v
            s = s.parent
        raise JSError(f"{name} is not defined")


# ---- array methods ------------------------------------------------------------------------------------------------

def _arr_sort(this, args):
    import functools
    cmp = args[0] if args else None
    if cmp is None:
        this.items.sort(key=to_string)
    else:
        def c(a, b):
            r = to_number(call(cmp, UNDEF, [a, b]))
            return -1 if r < 0 else (1 if r > 0 else 0)
        this.items.sort(key=functools.cmp_to_key(c))   # list.sort is stable, like Array.prototype.sort
    return this


def _arr_foreach(this, args):
    fn = args[0]
    for i, x in enumerate(list(this.items)):
        call(fn, UNDEF, [x, float(i), this])
    return UNDEF


def _arr_join(this, args):
    sep = to_string(args[0]) if args else ","
    return sep.join("" if x is UNDEF or x is None else to_string(x) for x in this.items)


def _arr_push(this, args):
    this.items.extend(args)
    return float(len(this.items))


_ARRAY_METHODS = {"sort": _arr_sort, "forEach": _arr_foreach, "join": _arr_join, "push": _arr_push}

# ---- tokenizer ------------------------------------------------------------------------------------------------------

_TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>0[xX][0-9a-fA-F]+|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)
  | (?P<id>[A-Za-z_$][A-Za-z0-9_$]*)
  | (?P<str>'(?:[^'\\]|\\.)*'|"(?:[^"\\]|\\.)*")
  | (?P<op>\.\.\.|===|!==|>>>|<<=|>>=|==|!=|<=|>=|&&|\|\||\+\+|--|\+=|-=|\*=|/=|%=|&=|\|=|\^=|<<|>>|=>|[-+*/%<>=!&|^~?:;,.(){}\[\]])
""", re.X | re.S)

_KEYWORDS = {"function", "return", "if", "else", "for", "while", "do", "break", "continue", "const", "let", "var",
             "new", "typeof", "true", "false", "null", "undefined", "this", "in", "of", "instanceof", "void", "delete"}


def tokenize(src):
    out, pos = [], 0
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise SyntaxError(f"bad character at {pos}: {src[pos:pos + 20]!r}")
        pos = m.end()
        k = m.lastgroup
        if k == "ws":
            continue
        t = m.group(k)
        if k == "num":
            out.append(("num", float(int(t, 16)) if t[:2] in ("0x", "0X") else float(t)))
        elif k == "str":
            out.append(("str", bytes(t[1:-1], "utf-8").decode("unicode_escape") if "\\" in t else t[1:-1]))
        elif k == "id":
            out.append(("kw", t) if t in _KEYWORDS else ("id", t))
        else:
            out.append(("op", t))
    out.append(("eof", None))
    return out


# ---- parser -> closures -------------------------------------------------------------------------------------------------

_BINPREC = {"||": 1, "&&": 2, "|": 3, "^": 4, "&": 5, "==": 6, "!=": 6, "===": 6, "!==": 6,
            "<": 7, ">": 7, "<=": 7, ">=": 7, "<<": 8, ">>": 8, ">>>": 8, "+": 9, "-": 9, "*": 10, "/": 10, "%": 10}


class Parser:
    def __init__(self, src):
        self.t = tokenize(src)
        self.i = 0
        self.hoist_stack = [[]]

    # -- token helpers
    def peek(self, k=0):
        return self.t[self.i + k]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def is_op(self, v):
        tok = self.t[self.i]
        return tok[0] == "op" and tok[1] == v

    def is_kw(self, v):
        tok = self.t[self.i]
        return tok[0] == "kw" and tok[1] == v

    def eat_op(self, v):
        if self.is_op(v):
            self.i += 1
            return True
        return False

    def expect_op(self, v):
        if not self.eat_op(v):
            raise SyntaxError(f"expected {v!r}, got {self.peek()!r} at token {self.i}")

    def ident(self):
        tok = self.next()
        if tok[0] != "id":
            raise SyntaxError(f"expected identifier, got {tok!r}")
        return tok[1]

    # -- program / statements
    def program(self):
        stmts = []
        while self.peek()[0] != "eof":
            stmts.append(self.statement())
        return self.block_fn(stmts), self.hoist_stack[0]

    @staticmethod
    def block_fn(stmts):
        funs = [s for s in stmts if getattr(s, "is_fundecl", False)]
        rest = [s for s in stmts if not getattr(s, "is_fundecl", False)]
        order = funs + rest                              # function declarations are hoisted

        def run(sc):
            for s in order:
                s(sc)
        return run

    def statement(self):
        tok = self.peek()
        if tok[0] == "op" and tok[1] == "{":
            self.next()
            stmts = []
            while not self.is_op("}"):
                stmts.append(self.statement())
            self.next()
            return self.block_fn(stmts)
        if tok[0] == "op" and tok[1] == ";":
            self.next()
            return lambda sc: None
        if tok[0] == "kw":
            k = tok[1]
            if k in ("const", "let", "var"):
                self.next()
                s = self.var_decl()
                self.eat_op(";")
                return s
            if k == "function" and self.peek(1)[0] == "id":
                return self.function_decl()
            if k == "return":
                self.next()
                if self.is_op(";") or self.is_op("}"):
                    self.eat_op(";")

                    def ret0(sc):
                        raise _Return(UNDEF)
                    return ret0
                e = self.expression()
                self.eat_op(";")

                def ret(sc):
                    raise _Return(e(sc))
                return ret
            if k == "if":
                self.next()
                self.expect_op("(")
                c = self.expression()
                self.expect_op(")")
                a = self.statement()
                b = None
                if self.is_kw("else"):
                    self.next()
                    b = self.statement()
                if b is None:
                    def if1(sc):
                        if truthy(c(sc)):
                            a(sc)
                    return if1

                def if2(sc):
                    if truthy(c(sc)):
                        a(sc)
                    else:
                        b(sc)
                return if2
            if k == "for":
                return self.for_stmt()
            if k == "while":
                self.next()
                self.expect_op("(")
                c = self.expression()
                self.expect_op(")")
                body = self.statement()

                def wh(sc):
                    while truthy(c(sc)):
                        try:
                            body(sc)
                        except _Break:
                            break
                        except _Continue:
                            continue
                return wh
            if k == "break":
                self.next()
                self.eat_op(";")

                def br(sc):
                    raise _Break()
                return br
            if k == "continue":
                self.next()
                self.eat_op(";")

                def co(sc):
                    raise _Continue()
                return co
        if tok[0] == "str" and self.peek(1) == ("op", ";"):      # directive prologue ('use strict')
            self.next()
            self.next()
            return lambda sc: None
        e = self.expression()
        self.eat_op(";")
        return e

    def var_decl(self):
        decls = []
        while True:
            name = self.ident()
            self.hoist_stack[-1].append(name)
            init = None
            if self.eat_op("="):
                init = self.assignment()
            decls.append((name, init))
            if not self.eat_op(","):
                break

        def run(sc):
            for name, init in decls:
                if init is not None:
                    sc.vars[name] = init(sc)
        return run

    def for_stmt(self):
        self.next()
        self.expect_op("(")
        init = None
        if not self.is_op(";"):
            if self.peek()[0] == "kw" and self.peek()[1] in ("let", "const", "var"):
                self.next()
                init = self.var_decl()
            else:
                init = self.expression()
        self.expect_op(";")
        cond = None if self.is_op(";") else self.expression()
        self.expect_op(";")
        upd = None if self.is_op(")") else self.expression()
        self.expect_op(")")
        body = self.statement()

        def run(sc):
            if init is not None:
                init(sc)
            while cond is None or truthy(cond(sc)):
                try:
                    body(sc)
                except _Break:
                    break
                except _Continue:
                    pass
                if upd is not None:
                    upd(sc)
        return run

    def function_decl(self):
        self.next()
        name = self.ident()
        self.hoist_stack[-1].append(name)
        make = self.function_rest(name)

        def run(sc):
            sc.vars[name] = make(sc)
        run.is_fundecl = True
        return run

    def function_rest(self, name):
        self.expect_op("(")
        params, rest = [], None
        while not self.is_op(")"):
            if self.eat_op("..."):
                rest = self.ident()
            else:
                params.append(self.ident())
            if not self.eat_op(","):
                break
        self.expect_op(")")
        self.expect_op("{")
        self.hoist_stack.append([])
        stmts = []
        while not self.is_op("}"):
            stmts.append(self.statement())
        self.next()
        hoist = tuple(dict.fromkeys(self.hoist_stack.pop()))
        body = self.block_fn(stmts)
        return lambda sc: JSFunction(params, rest, body, sc, name, hoist)

    # -- expressions
    def expression(self):
        e = self.assignment()
        while self.is_op(","):
            self.next()
            r = self.assignment()
            l = e
            e = (lambda l, r: lambda sc: (l(sc), r(sc))[1])(l, r)
        return e

    def assignment(self):
        start = self.i
        left = self.conditional()
        tok = self.peek()
        if tok[0] == "op" and tok[1] in ("=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>="):
            ref = getattr(left, "ref", None)
            if ref is None:
                raise SyntaxError(f"invalid assignment target at token {start}")
            self.next()
            right = self.assignment()
            op = tok[1]
            if op == "=":
                if ref[0] == "var":
                    name = ref[1]

                    def asg(sc):
                        v = right(sc)
                        sc.lookup(name)[name] = v
                        return v
                    return asg
                objf, keyf = ref[1], ref[2]

                def asgm(sc):
                    o = objf(sc)
                    k = keyf(sc)
                    v = right(sc)
                    o.js_set(k, v)
                    return v
                return asgm
            binop = _binop(op[:-1])
            if ref[0] == "var":
                name = ref[1]

                def casg(sc):
                    d = sc.lookup(name)
                    v = binop(d[name], right(sc))
                    d[name] = v
                    return v
                return casg
            objf, keyf = ref[1], ref[2]

            def casgm(sc):
                o = objf(sc)
                k = keyf(sc)
                v = binop(get_member(o, k), right(sc))
                o.js_set(k, v)
                return v
            return casgm
        return left

    def conditional(self):
        c = self.binary(0)
        if self.eat_op("?"):
            a = self.assignment()
            self.expect_op(":")
            b = self.assignment()
            return lambda sc: a(sc) if truthy(c(sc)) else b(sc)
        return c

    def binary(self, minprec):
        left = self.unary()
        while True:
            tok = self.peek()
            if tok[0] != "op" or tok[1] not in _BINPREC or _BINPREC[tok[1]] <= minprec - 1 and False:
                break
            prec = _BINPREC[tok[1]]
            if prec <= minprec:
                break
            self.next()
            right = self.binary(prec)
            op = tok[1]
            l = left
            if op == "&&":
                left = (lambda l, r: lambda sc: (lambda v: r(sc) if truthy(v) else v)(l(sc)))(l, right)
            elif op == "||":
                left = (lambda l, r: lambda sc: (lambda v: v if truthy(v) else r(sc))(l(sc)))(l, right)
            else:
                f = _binop(op)
                left = (lambda l, r, f: lambda sc: f(l(sc), r(sc)))(l, right, f)
        return left

    def unary(self):
        tok = self.peek()
        if tok[0] == "op":
            if tok[1] == "-":
                self.next()
                e = self.unary()
                return lambda sc: -to_number(e(sc))
            if tok[1] == "+":
                self.next()
                e = self.unary()
                return lambda sc: to_number(e(sc))
            if tok[1] == "!":
                self.next()
                e = self.unary()
                return lambda sc: not truthy(e(sc))
            if tok[1] == "~":
                self.next()
                e = self.unary()
                return lambda sc: float(~to_int32(e(sc)))
            if tok[1] in ("++", "--"):
                self.next()
                e = self.unary()
                return self.update(e, 1.0 if tok[1] == "++" else -1.0, prefix=True)
        if tok[0] == "kw" and tok[1] == "typeof":
            self.next()
            e = self.unary()
            return lambda sc: js_typeof(e(sc))
        return self.postfix()

    def update(self, e, delta, prefix):
        ref = getattr(e, "ref", None)
        if ref is None:
            raise SyntaxError("invalid update target")
        if ref[0] == "var":
            name = ref[1]

            def upv(sc):
                d = sc.lookup(name)
                old = to_number(d[name])
                d[name] = old + delta
                return old + delta if prefix else old
            return upv
        objf, keyf = ref[1], ref[2]

        def upm(sc):
            o = objf(sc)
            k = keyf(sc)
            old = to_number(get_member(o, k))
            o.js_set(k, old + delta)
            return old + delta if prefix else old
        return upm

    def postfix(self):
        e = self.call_member()
        tok = self.peek()
        if tok[0] == "op" and tok[1] in ("++", "--"):
            self.next()
            return self.update(e, 1.0 if tok[1] == "++" else -1.0, prefix=False)
        return e

    def arguments(self):
        args = []
        self.expect_op("(")
        while not self.is_op(")"):
            args.append(self.assignment())
            if not self.eat_op(","):
                break
        self.expect_op(")")
        return args

    def call_member(self):
        tok = self.peek()
        if tok[0] == "kw" and tok[1] == "new":
            self.next()
            ctor = self.member_only()
            args = self.arguments() if self.is_op("(") else []
            e = (lambda ctor, args: lambda sc: construct(ctor(sc), [a(sc) for a in args]))(ctor, args)
        else:
            e = self.primary()
        while True:
            if self.eat_op("."):
                name = self.next()[1]
                obj = e
                e = (lambda obj, name: lambda sc: get_member(obj(sc), name))(obj, name)
                e.ref = ("member", obj, (lambda name: lambda sc: name)(name))
                e.method = (obj, (lambda name: lambda sc: name)(name))
            elif self.is_op("["):
                self.next()
                key = self.expression()
                self.expect_op("]")
                obj = e
                e = (lambda obj, key: lambda sc: get_member(obj(sc), key(sc)))(obj, key)
                e.ref = ("member", obj, key)
                e.method = (obj, key)
            elif self.is_op("("):
                args = self.arguments()
                method = getattr(e, "method", None)
                if method is not None:
                    objf, keyf = method

                    def mcall(sc, objf=objf, keyf=keyf, args=args):
                        o = objf(sc)
                        fn = get_member(o, keyf(sc))
                        return call(fn, o, [a(sc) for a in args])
                    e = mcall
                else:
                    fnf = e
                    e = (lambda fnf, args: lambda sc: call(fnf(sc), UNDEF, [a(sc) for a in args]))(fnf, args)
            else:
                return e

    def member_only(self):
        e = self.primary()
        while self.eat_op("."):
            name = self.next()[1]
            obj = e
            e = (lambda obj, name: lambda sc: get_member(obj(sc), name))(obj, name)
        return e

    def primary(self):
        tok = self.next()
        k, v = tok
        if k == "num":
            return lambda sc: v
        if k == "str":
            return lambda sc: v
        if k == "id":
            def var(sc):
                s = sc
                while s is not None:
                    d = s.vars
                    if v in d:
                        return d[v]
                    s = s.parent
                raise JSError(f"{v} is not defined")
            var.ref = ("var", v)
            return var
        if k == "kw":
            if v == "true":
                return lambda sc: True
            if v == "false":
                return lambda sc: False
            if v == "null":
                return lambda sc: None
            if v == "undefined":
                return lambda sc: UNDEF
            if v == "this":
                def this(sc):
                    s = sc
                    while s is not None:
                        if "this" in s.vars:
                            return s.vars["this"]
                        s = s.parent
                    return UNDEF
                return this
            if v == "function":
                name = self.ident() if self.peek()[0] == "id" else ""
                return self.function_rest(name)
        if k == "op":
            if v == "(":
                e = self.expression()
                self.expect_op(")")
                return e
            if v == "[":
                items = []
                while not self.is_op("]"):
                    items.append(self.assignment())
                    if not self.eat_op(","):
                        break
                self.expect_op("]")
                return lambda sc: JSArray([f(sc) for f in items])
            if v == "{":
                props = []
                while not self.is_op("}"):
                    kt = self.next()
                    key = to_string(kt[1]) if kt[0] == "num" else kt[1]
                    self.expect_op(":")
                    props.append((key, self.assignment()))
                    if not self.eat_op(","):
                        break
                self.expect_op("}")
                return lambda sc: JSObject({key: f(sc) for key, f in props})
        raise SyntaxError(f"unexpected token {tok!r} at {self.i}")


def _cmp(op):
    def f(a, b):
        if isinstance(a, str) and isinstance(b, str):
            return op(a, b)
        return op(to_number(a), to_number(b))     # comparisons with NaN are false in Python too
    return f


def _binop(op):
    import operator as o
    if op == "+":
        return js_add
    if op == "-":
        return lambda a, b: to_number(a) - to_number(b)
    if op == "*":
        return lambda a, b: to_number(a) * to_number(b)
    if op == "/":
        return js_div
    if op == "%":
        return lambda a, b: math.fmod(to_number(a), to_number(b)) if to_number(b) != 0 else math.nan
    if op == "<":
        return _cmp(o.lt)
    if op == ">":
        return _cmp(o.gt)
    if op == "<=":
        return _cmp(o.le)
    if op == ">=":
        return _cmp(o.ge)
    if op == "==":
        return loose_eq
    if op == "!=":
        return lambda a, b: not loose_eq(a, b)
    if op == "===":
        return strict_eq
    if op == "!==":
        return lambda a, b: not strict_eq(a, b)
    if op == "&":
        return lambda a, b: float(to_int32(a) & to_int32(b))
    if op == "|":
        return lambda a, b: float(to_int32(a) | to_int32(b))
    if op == "^":
        return lambda a, b: float(to_int32(a) ^ to_int32(b))
    if op == "<<":
        return lambda a, b: float(to_int32(to_int32(a) << (to_int32(b) & 31)))
    if op == ">>":
        return lambda a, b: float(to_int32(a) >> (to_int32(b) & 31))
    raise SyntaxError(f"operator {op} not supported")


# ---- standard library the program needs -------------------------------------------------------------------------------

def _safe(fn):
    def g(*a):
        try:
            return float(fn(*a))
        except (ValueError, OverflowError):
            return math.nan
    return g


def _js_pow(x, y):
    try:
        return math.pow(x, y)
    except OverflowError:
        return math.inf
    except ValueError:
        return math.nan


def _js_ceil(x):
    if x != x or x in (math.inf, -math.inf):
        return x
    return float(math.ceil(x))


def _js_max(args):
    r = -math.inf
    for a in args:
        a = to_number(a)
        if a != a:
            return math.nan
        if a > r or (a == 0 and r == 0 and math.copysign(1, r) < 0):
            r = a
    return r


def _js_min(args):
    r = math.inf
    for a in args:
        a = to_number(a)
        if a != a:
            return math.nan
        if a < r or (a == 0 and r == 0 and math.copysign(1, a) < 0):
            r = a
    return r


def make_math(random_fn):
    nf = NativeFunction
    m = JSObject({
        "PI": math.pi,
        "sqrt": nf(lambda t, a: _safe(math.sqrt)(to_number(a[0])), "sqrt"),
        "tan": nf(lambda t, a: _safe(math.tan)(to_number(a[0])), "tan"),
        "atan2": nf(lambda t, a: _safe(math.atan2)(to_number(a[0]), to_number(a[1])), "atan2"),
        "asin": nf(lambda t, a: _safe(math.asin)(to_number(a[0])), "asin"),
        "pow": nf(lambda t, a: _js_pow(to_number(a[0]), to_number(a[1])), "pow"),
        "ceil": nf(lambda t, a: _js_ceil(to_number(a[0])), "ceil"),
        "floor": nf(lambda t, a: float(math.floor(to_number(a[0]))), "floor"),
        "abs": nf(lambda t, a: abs(to_number(a[0])), "abs"),
        "max": nf(lambda t, a: _js_max(a), "max"),
        "min": nf(lambda t, a: _js_min(a), "min"),
        "random": nf(lambda t, a: random_fn(), "random"),
    })
    return m


def json_stringify(v):
    import json

    def conv(x):
        if isinstance(x, JSArray):
            return [conv(i) for i in x.items]
        if isinstance(x, JSObject):
            return {k: conv(i) for k, i in x.props.items() if not isinstance(i, (JSFunction, NativeFunction))}
        if x is UNDEF:
            return None
        return x
    return json.dumps(conv(v), separators=(",", ":"))


class Interpreter:
    def __init__(self, random_fn=lambda: 0.5):
        self.timers = []                              # FIFO: setTimeout drained synchronously
        self.log = []
        g = {
            "Infinity": math.inf, "NaN": math.nan, "undefined": UNDEF,
            "Math": make_math(random_fn),
            "JSON": JSObject({"stringify": NativeFunction(lambda t, a: json_stringify(a[0]), "stringify")}),
            "Date": JSObject({"now": NativeFunction(lambda t, a: 0.0, "now")}),
            "setTimeout": NativeFunction(self._set_timeout, "setTimeout"),
            "console": JSObject({"log": NativeFunction(self._console_log, "log")}),
        }
        self.globals = Scope(g, None)

    def _set_timeout(self, this, args):
        self.timers.append((args[0], list(args[2:])))
        return float(len(self.timers))

    def _console_log(self, this, args):
        self.log.append(" ".join(to_string(a) for a in args))
        return UNDEF

    def run(self, src):
        body, hoist = Parser(src).program()
        for name in hoist:
            self.globals.vars.setdefault(name, UNDEF)
        body(self.globals)

    def drain_timers(self, max_calls=10 ** 9):
        n = 0
        while self.timers and n < max_calls:
            fn, args = self.timers.pop(0)
            call(fn, UNDEF, args)
            n += 1
        return n

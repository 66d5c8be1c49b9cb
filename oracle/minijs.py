"""minijs.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A small tree-walking interpreter for the ES5 subset the reference's hot-path
sources are written in (lib/jsfft/*.js, src/utils.js, src/extractors/*.js and the
method bodies of src/meyda.js).  No JavaScript engine exists in this image, so
this is how the reference's OWN source text gets executed here: the golden
vectors under tests/golden/js_reference_vectors.npz are what those unmodified
files compute under this interpreter (tools/make_js_golden.py), and the oracle
is pinned against them (tests/test_js_pin.py).

Semantics kept: every Number is an IEEE double; Float32Array / Int32Array stores
round as the typed arrays do; bitwise operators use ToInt32; property keys are
strings ("3", not "3.0"); prototype chains, `new`, `this`, closures, `typeof`,
`instanceof`, `hasOwnProperty`, `Function.prototype.bind/call/apply`,
`Array.apply(null, new Array(n)).map(Number.prototype.valueOf, 0)`; automatic
semicolon insertion at line breaks (jsfft is written without semicolons).
Not supported: getters/setters, regex literals, labels, `with`, `switch`,
generators, ES6 syntax (src/meyda.js's class is not executed as such; its method
bodies are lifted into plain functions by make_js_golden.py).
"""
from __future__ import annotations

import math
import os
import re

import numpy as np


class JSUndefined:
    _inst = None

    def __new__(cls):
        if cls._inst is None:
            cls._inst = super().__new__(cls)
        return cls._inst

    def __repr__(self):
        return "undefined"

    def __bool__(self):
        return False


undefined = JSUndefined()


class JSError(Exception):
    def __init__(self, value):
        super().__init__(str(value))
        self.value = value


class JSObject:
    def __init__(self, proto=None):
        self.props = {}
        self.proto = proto

    def get(self, key):
        o = self
        while o is not None:
            if key in o.props:
                return o.props[key]
            o = o.proto
        return undefined

    def put(self, key, value):
        self.props[key] = value

    def has_own(self, key):
        return key in self.props


class JSFunction(JSObject):
    def __init__(self, interp, params, body, env, name="", native=None):
        super().__init__(interp.function_proto if interp is not None else None)
        self.interp, self.params, self.body, self.env, self.name, self.native = interp, params, body, env, name, native
        self.bound_this = None
        self.bound_target = None
        if native is None and interp is not None:
            p = JSObject(interp.object_proto)
            p.props["constructor"] = self
            self.props["prototype"] = p

    def call(self, this, args):
        if self.bound_target is not None:
            return self.bound_target.call(self.bound_this, list(self.bound_args) + list(args))
        if self.native is not None:
            return self.native(this, args)
        return self.interp.call_function(self, this, args)


class JSArray(JSObject):
    def __init__(self, interp, items=None):
        super().__init__(interp.array_proto)
        self.items = list(items) if items is not None else []

    def get(self, key):
        if key == "length":
            return float(len(self.items))
        i = _array_index(key)
        if i is not None:
            return self.items[i] if i < len(self.items) else undefined
        return super().get(key)

    def put(self, key, value):
        if key == "length":
            n = int(value)
            del self.items[n:]
            self.items.extend([undefined] * (n - len(self.items)))
            return
        i = _array_index(key)
        if i is not None:
            if i >= len(self.items):
                self.items.extend([undefined] * (i + 1 - len(self.items)))
            self.items[i] = value
        else:
            super().put(key, value)

    def has_own(self, key):
        i = _array_index(key)
        if i is not None:
            return i < len(self.items) and self.items[i] is not undefined
        return key == "length" or super().has_own(key)


class JSTypedArray(JSObject):
    def __init__(self, interp, kind, data):
        super().__init__(interp.typed_protos[kind])
        self.kind, self.data = kind, data

    def get(self, key):
        if key == "length":
            return float(len(self.data))
        i = _array_index(key)
        if i is not None:
            return float(self.data[i]) if i < len(self.data) else undefined
        return super().get(key)

    def put(self, key, value):
        i = _array_index(key)
        if i is not None:
            if i < len(self.data):  # out-of-range typed-array stores are ignored
                v = to_number(value)
                with np.errstate(all="ignore"):
                    if self.kind == "Float32Array":
                        self.data[i] = np.float32(v)
                    elif self.kind == "Float64Array":
                        self.data[i] = v
                    else:
                        self.data[i] = to_int32(v)
        else:
            super().put(key, value)

    def has_own(self, key):
        i = _array_index(key)
        return (i is not None and i < len(self.data)) or key == "length" or super().has_own(key)


def _array_index(key):
    if isinstance(key, str) and key.isdigit() and (key == "0" or key[0] != "0"):
        return int(key)
    return None


def to_number(v):
    if isinstance(v, bool):
        return 1.0 if v else 0.0
    if isinstance(v, (int, float)):
        return float(v)
    if v is undefined:
        return math.nan
    if v is None:
        return 0.0
    if isinstance(v, str):
        s = v.strip()
        if s == "":
            return 0.0
        try:
            return float(s)
        except ValueError:
            return math.nan
    if isinstance(v, np.floating):
        return float(v)
    return math.nan


def to_int32(v):
    v = to_number(v)
    if v != v or math.isinf(v):
        return 0
    n = int(math.copysign(math.floor(abs(v)), v)) & 0xFFFFFFFF
    return n - 0x100000000 if n >= 0x80000000 else n


def to_uint32(v):
    return to_int32(v) & 0xFFFFFFFF


def to_boolean(v):
    if isinstance(v, bool):
        return v
    if isinstance(v, (int, float)):
        return not (v == 0 or v != v)
    if v is undefined or v is None:
        return False
    if isinstance(v, str):
        return len(v) > 0
    return True


def num_to_str(x):
    if x != x:
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == math.floor(x) and abs(x) < 1e21:
        return str(int(x))
    return repr(x)


def to_str(v):
    if isinstance(v, str):
        return v
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, float)):
        return num_to_str(float(v))
    if v is undefined:
        return "undefined"
    if v is None:
        return "null"
    if isinstance(v, JSFunction):
        return "function"
    if isinstance(v, JSArray):
        return ",".join("" if (x is undefined or x is None) else to_str(x) for x in v.items)
    return "[object Object]"


def to_key(v):
    return to_str(v)


def js_typeof(v):
    if v is undefined:
        return "undefined"
    if v is None:
        return "object"
    if isinstance(v, bool):
        return "boolean"
    if isinstance(v, (int, float)):
        return "number"
    if isinstance(v, str):
        return "string"
    if isinstance(v, JSFunction):
        return "function"
    return "object"


# ------------------------------------------------------------------ tokenizer
_PUNCT = [">>>=", "===", "!==", ">>>", "<<=", ">>=", "&&", "||", "==", "!=", "<=", ">=", "++", "--", "+=", "-=", "*=",
          "/=", "%=", "&=", "|=", "^=", "<<", ">>", "{", "}", "(", ")", "[", "]", ";", ",", "<", ">", "+", "-", "*", "/",
          "%", "&", "|", "^", "!", "~", "?", ":", "=", "."]
_KEYWORDS = {"var", "function", "return", "if", "else", "for", "while", "do", "break", "continue", "new", "this",
             "typeof", "instanceof", "in", "null", "true", "false", "throw", "try", "catch", "finally", "delete",
             "void", "undefined_kw_never"}
_ID_START = re.compile(r"[A-Za-z_$ª-￿]")
_ID_PART = re.compile(r"[A-Za-z0-9_$ª-￿]")
_NUM = re.compile(r"0[xX][0-9a-fA-F]+|(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?")


class Tok:
    __slots__ = ("kind", "value", "nl", "pos")

    def __init__(self, kind, value, nl, pos):
        self.kind, self.value, self.nl, self.pos = kind, value, nl, pos

    def __repr__(self):
        return "%s(%r)" % (self.kind, self.value)


def tokenize(src):
    toks, i, n, nl = [], 0, len(src), False
    while i < n:
        c = src[i]
        if c in " \t\r\ufeff\u00a0\u000b\u000c":
            i += 1
        elif c == "\n":
            nl = True
            i += 1
        elif src.startswith("//", i):
            while i < n and src[i] != "\n":
                i += 1
        elif src.startswith("/*", i):
            j = src.index("*/", i + 2)
            if "\n" in src[i:j]:
                nl = True
            i = j + 2
        elif c in "'\"":
            j, out = i + 1, []
            while src[j] != c:
                if src[j] == "\\":
                    j += 1
                    out.append({"n": "\n", "t": "\t", "r": "\r", "0": "\0"}.get(src[j], src[j]))
                else:
                    out.append(src[j])
                j += 1
            toks.append(Tok("str", "".join(out), nl, i))
            nl = False
            i = j + 1
        elif c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit()):
            m = _NUM.match(src, i)
            t = m.group(0)
            toks.append(Tok("num", float(int(t, 16)) if t[:2].lower() == "0x" else float(t), nl, i))
            nl = False
            i = m.end()
        elif _ID_START.match(c):
            j = i + 1
            while j < n and _ID_PART.match(src[j]):
                j += 1
            w = src[i:j]
            toks.append(Tok("kw" if w in _KEYWORDS else "id", w, nl, i))
            nl = False
            i = j
        else:
            for p in _PUNCT:
                if src.startswith(p, i):
                    toks.append(Tok("p", p, nl, i))
                    nl = False
                    i += len(p)
                    break
            else:
                raise SyntaxError("unexpected character %r at %d" % (c, i))
    toks.append(Tok("eof", None, True, n))
    return toks


# --------------------------------------------------------------------- parser
_BINARY = [("||",), ("&&",), ("|",), ("^",), ("&",), ("==", "!=", "===", "!=="),
           ("<", ">", "<=", ">=", "instanceof", "in"), ("<<", ">>", ">>>"), ("+", "-"), ("*", "/", "%")]
_ASSIGN = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>=", ">>>="}


class Parser:
    def __init__(self, src):
        self.t = tokenize(src)
        self.i = 0
        self.no_in = False

    def peek(self):
        return self.t[self.i]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def is_p(self, v):
        tok = self.t[self.i]
        return tok.kind == "p" and tok.value == v

    def is_kw(self, v):
        tok = self.t[self.i]
        return tok.kind == "kw" and tok.value == v

    def eat_p(self, v):
        if self.is_p(v):
            self.i += 1
            return True
        return False

    def expect_p(self, v):
        if not self.eat_p(v):
            raise SyntaxError("expected %r, got %r at %d" % (v, self.peek(), self.peek().pos))

    def end_statement(self):
        if self.eat_p(";"):
            return
        tok = self.peek()
        if tok.kind == "eof" or tok.nl or (tok.kind == "p" and tok.value == "}"):
            return  # automatic semicolon insertion
        raise SyntaxError("expected ';', got %r at %d" % (tok, tok.pos))

    def program(self):
        body = []
        while self.peek().kind != "eof":
            body.append(self.statement())
        return ("block", body)

    def block(self):
        self.expect_p("{")
        body = []
        while not self.is_p("}"):
            body.append(self.statement())
        self.expect_p("}")
        return ("block", body)

    def statement(self):
        tok = self.peek()
        if tok.kind == "p":
            if tok.value == "{":
                return self.block()
            if tok.value == ";":
                self.next()
                return ("empty",)
        if tok.kind == "kw":
            k = tok.value
            if k == "var":
                self.next()
                d = self.var_decls()
                self.end_statement()
                return d
            if k == "function":
                self.next()
                name = self.next().value
                return ("funcdecl", name, self.function_rest(name))
            if k == "if":
                self.next()
                self.expect_p("(")
                c = self.expression()
                self.expect_p(")")
                a = self.statement()
                b = None
                if self.is_kw("else"):
                    self.next()
                    b = self.statement()
                return ("if", c, a, b)
            if k == "for":
                return self.for_statement()
            if k == "while":
                self.next()
                self.expect_p("(")
                c = self.expression()
                self.expect_p(")")
                return ("while", c, self.statement())
            if k == "do":
                self.next()
                body = self.statement()
                if not self.is_kw("while"):
                    raise SyntaxError("expected while")
                self.next()
                self.expect_p("(")
                c = self.expression()
                self.expect_p(")")
                self.eat_p(";")
                return ("dowhile", c, body)
            if k == "return":
                self.next()
                tok2 = self.peek()
                val = None
                if not (tok2.nl or tok2.kind == "eof" or (tok2.kind == "p" and tok2.value in (";", "}"))):
                    val = self.expression()
                self.end_statement()
                return ("return", val)
            if k in ("break", "continue"):
                self.next()
                self.end_statement()
                return (k,)
            if k == "throw":
                self.next()
                e = self.expression()
                self.end_statement()
                return ("throw", e)
            if k == "try":
                self.next()
                body = self.block()
                param = handler = final = None
                if self.is_kw("catch"):
                    self.next()
                    self.expect_p("(")
                    param = self.next().value
                    self.expect_p(")")
                    handler = self.block()
                if self.is_kw("finally"):
                    self.next()
                    final = self.block()
                return ("try", body, param, handler, final)
        e = self.expression()
        self.end_statement()
        return ("expr", e)

    def var_decls(self):
        decls = []
        while True:
            name = self.next()
            if name.kind != "id":
                raise SyntaxError("bad var name %r" % (name,))
            init = None
            if self.eat_p("="):
                init = self.assignment()
            decls.append((name.value, init))
            if not self.eat_p(","):
                break
        return ("var", decls)

    def for_statement(self):
        self.next()
        self.expect_p("(")
        init = None
        if self.is_kw("var"):
            self.next()
            self.no_in = True
            init = self.var_decls()
            self.no_in = False
            if self.is_kw("in"):
                self.next()
                obj = self.expression()
                self.expect_p(")")
                return ("forin", init[1][0][0], obj, self.statement())
        elif not self.is_p(";"):
            self.no_in = True
            init = ("expr", self.expression())
            self.no_in = False
        self.expect_p(";")
        test = None if self.is_p(";") else self.expression()
        self.expect_p(";")
        update = None if self.is_p(")") else self.expression()
        self.expect_p(")")
        return ("for", init, test, update, self.statement())

    def function_rest(self, name):
        self.expect_p("(")
        params = []
        while not self.is_p(")"):
            params.append(self.next().value)
            self.eat_p(",")
        self.expect_p(")")
        return ("function", name, params, self.block())

    def expression(self):
        e = self.assignment()
        while self.is_p(","):
            self.next()
            e = ("comma", e, self.assignment())
        return e

    def assignment(self):
        left = self.conditional()
        tok = self.peek()
        if tok.kind == "p" and tok.value in _ASSIGN:
            self.next()
            return ("assign", tok.value, left, self.assignment())
        return left

    def conditional(self):
        c = self.binary(0)
        if self.is_p("?"):
            self.next()
            a = self.assignment()
            self.expect_p(":")
            return ("cond", c, a, self.assignment())
        return c

    def binary(self, level):
        if level == len(_BINARY):
            return self.unary()
        left = self.binary(level + 1)
        while True:
            tok = self.peek()
            if tok.kind in ("p", "kw") and tok.value in _BINARY[level] and not (tok.value == "in" and self.no_in):
                self.next()
                left = ("bin", tok.value, left, self.binary(level + 1))
            else:
                return left

    def unary(self):
        tok = self.peek()
        if tok.kind == "p" and tok.value in ("!", "-", "+", "~"):
            self.next()
            return ("unary", tok.value, self.unary())
        if tok.kind == "p" and tok.value in ("++", "--"):
            self.next()
            return ("update", tok.value, True, self.unary())
        if tok.kind == "kw" and tok.value in ("typeof", "void", "delete"):
            self.next()
            return ("unary", tok.value, self.unary())
        e = self.postfix_calls()
        tok = self.peek()
        if tok.kind == "p" and tok.value in ("++", "--") and not tok.nl:
            self.next()
            return ("update", tok.value, False, e)
        return e

    def postfix_calls(self):
        if self.is_kw("new"):
            self.next()
            callee = self.member_only()
            args = self.arguments() if self.is_p("(") else []
            e = ("new", callee, args)
        else:
            e = self.primary()
        while True:
            if self.is_p("."):
                self.next()
                e = ("member", e, ("str", self.next().value))
            elif self.is_p("["):
                self.next()
                k = self.expression()
                self.expect_p("]")
                e = ("member", e, k)
            elif self.is_p("("):
                e = ("call", e, self.arguments())
            else:
                return e

    def member_only(self):
        if self.is_kw("new"):
            self.next()
            callee = self.member_only()
            args = self.arguments() if self.is_p("(") else []
            return ("new", callee, args)
        e = self.primary()
        while True:
            if self.is_p("."):
                self.next()
                e = ("member", e, ("str", self.next().value))
            elif self.is_p("["):
                self.next()
                k = self.expression()
                self.expect_p("]")
                e = ("member", e, k)
            else:
                return e

    def arguments(self):
        self.expect_p("(")
        args = []
        while not self.is_p(")"):
            args.append(self.assignment())
            self.eat_p(",")
        self.expect_p(")")
        return args

    def primary(self):
        tok = self.next()
        if tok.kind == "num":
            return ("num", tok.value)
        if tok.kind == "str":
            return ("str", tok.value)
        if tok.kind == "id":
            return ("id", tok.value)
        if tok.kind == "kw":
            if tok.value == "this":
                return ("this",)
            if tok.value == "null":
                return ("null",)
            if tok.value == "true":
                return ("bool", True)
            if tok.value == "false":
                return ("bool", False)
            if tok.value == "function":
                name = ""
                if self.peek().kind == "id":
                    name = self.next().value
                return self.function_rest(name)
        if tok.kind == "p":
            if tok.value == "(":
                save = self.no_in
                self.no_in = False
                e = self.expression()
                self.no_in = save
                self.expect_p(")")
                return e
            if tok.value == "[":
                items = []
                while not self.is_p("]"):
                    items.append(self.assignment())
                    self.eat_p(",")
                self.expect_p("]")
                return ("array", items)
            if tok.value == "{":
                props = []
                while not self.is_p("}"):
                    k = self.next()
                    key = num_to_str(k.value) if k.kind == "num" else k.value
                    self.expect_p(":")
                    props.append((key, self.assignment()))
                    self.eat_p(",")
                self.expect_p("}")
                return ("object", props)
        raise SyntaxError("unexpected token %r at %d" % (tok, tok.pos))


# ---------------------------------------------------------------- interpreter
class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    def __init__(self, value):
        self.value = value


class Env:
    __slots__ = ("vars", "parent")

    def __init__(self, parent=None):
        self.vars = {}
        self.parent = parent

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e
            e = e.parent
        return None


def _hoist(node, out):
    """var / function declarations of a function body (not descending into nested functions)."""
    k = node[0]
    if k == "var":
        out.extend(n for n, _ in node[1])
    elif k == "funcdecl":
        out.append(node[1])
    elif k == "block":
        for s in node[1]:
            _hoist(s, out)
    elif k == "if":
        _hoist(node[2], out)
        if node[3]:
            _hoist(node[3], out)
    elif k == "for":
        if node[1]:
            _hoist(node[1], out)
        _hoist(node[4], out)
    elif k == "forin":
        out.append(node[1])
        _hoist(node[3], out)
    elif k in ("while", "dowhile"):
        _hoist(node[2], out)
    elif k == "try":
        _hoist(node[1], out)
        if node[3]:
            _hoist(node[3], out)
        if node[4]:
            _hoist(node[4], out)


class Interpreter:
    def __init__(self, root):
        self.root = root
        self.object_proto = JSObject(None)
        self.function_proto = JSObject(self.object_proto)
        self.array_proto = JSObject(self.object_proto)
        self.typed_protos = {k: JSObject(self.object_proto) for k in ("Float32Array", "Float64Array", "Int32Array")}
        self.global_env = Env()
        self.global_obj = JSObject(self.object_proto)
        self.modules = {}
        self._setup()

    # -- helpers
    def native(self, fn, name=""):
        f = JSFunction(self, [], None, None, name, native=fn)
        return f

    def _setup(self):
        G = self.global_env.vars
        op, fp, ap = self.object_proto, self.function_proto, self.array_proto
        op.props["hasOwnProperty"] = self.native(lambda this, a: self._has_own(this, a[0] if a else undefined))
        op.props["toString"] = self.native(lambda this, a: to_str(this))
        op.props["valueOf"] = self.native(lambda this, a: this)

        def f_call(this, a):
            return this.call(a[0] if a else undefined, list(a[1:]))

        def f_apply(this, a):
            arr = a[1] if len(a) > 1 else undefined
            items = self._array_like(arr)
            return this.call(a[0] if a else undefined, items)

        def f_bind(this, a):
            b = JSFunction(self, [], None, None, "bound")
            b.bound_target, b.bound_this, b.bound_args = this, (a[0] if a else undefined), list(a[1:])
            return b

        fp.props["call"] = self.native(f_call)
        fp.props["apply"] = self.native(f_apply)
        fp.props["bind"] = self.native(f_bind)

        def a_map(this, a):
            fn, this_arg = a[0], (a[1] if len(a) > 1 else undefined)
            out = JSArray(self)
            out.items = [undefined] * len(this.items)
            for i, v in enumerate(this.items):
                out.items[i] = fn.call(this_arg, [v, float(i), this])
            return out

        def a_push(this, a):
            this.items.extend(a)
            return float(len(this.items))

        def a_foreach(this, a):
            for i, v in enumerate(list(this.items)):
                a[0].call(a[1] if len(a) > 1 else undefined, [v, float(i), this])
            return undefined

        ap.props["map"] = self.native(a_map)
        ap.props["push"] = self.native(a_push)
        ap.props["forEach"] = self.native(a_foreach)
        ap.props["join"] = self.native(lambda this, a: (a[0] if a else ",").join(to_str(x) for x in this.items))

        def array_ctor(this, a):
            arr = JSArray(self)
            if len(a) == 1 and isinstance(a[0], (int, float)) and not isinstance(a[0], bool):
                arr.items = [undefined] * int(a[0])
            else:
                arr.items = list(a)
            return arr

        Array = self.native(array_ctor, "Array")
        Array.props["prototype"] = ap
        G["Array"] = Array

        def typed_ctor(kind, dtype):
            def ctor(this, a):
                x = a[0] if a else 0.0
                if isinstance(x, JSTypedArray):
                    data = x.data.astype(dtype)
                elif isinstance(x, JSArray):
                    data = np.array([to_number(v) for v in x.items], dtype=np.float64).astype(dtype)
                else:
                    n = to_number(x)
                    data = np.zeros(0 if n != n else int(n), dtype=dtype)
                return JSTypedArray(self, kind, data)
            f = self.native(ctor, kind)
            f.props["prototype"] = self.typed_protos[kind]

            def subarray(this, a):  # a view on the same storage, as in JS (numpy basic slicing)
                n = len(this.data)
                lo = int(to_number(a[0])) if a else 0
                hi = int(to_number(a[1])) if len(a) > 1 and a[1] is not undefined else n
                lo, hi = (max(n + v, 0) if v < 0 else min(v, n) for v in (lo, hi))
                return JSTypedArray(self, kind, this.data[lo:max(hi, lo)])
            self.typed_protos[kind].props["subarray"] = self.native(subarray, "subarray")
            return f

        G["Float32Array"] = typed_ctor("Float32Array", np.float32)
        G["Float64Array"] = typed_ctor("Float64Array", np.float64)
        G["Int32Array"] = typed_ctor("Int32Array", np.int32)

        Number = self.native(lambda this, a: to_number(a[0]) if a else 0.0, "Number")
        nproto = JSObject(op)
        nproto.props["valueOf"] = self.native(lambda this, a: this)
        nproto.props["toFixed"] = self.native(lambda this, a: ("%." + str(int(a[0]) if a else 0) + "f") % this)
        Number.props["prototype"] = nproto
        self.number_proto = nproto
        G["Number"] = Number
        Obj = self.native(lambda this, a: JSObject(op), "Object")
        Obj.props["prototype"] = op
        G["Object"] = Obj

        def err_ctor(this, a):
            o = this if isinstance(this, JSObject) and not isinstance(this, JSFunction) else JSObject(op)
            o.props["message"] = to_str(a[0]) if a else ""
            return o

        Err = self.native(err_ctor, "Error")
        Err.props["prototype"] = JSObject(op)
        G["Error"] = Err

        M = JSObject(op)
        M.props.update(PI=math.pi, SQRT1_2=math.sqrt(0.5), SQRT2=math.sqrt(2.0), E=math.e, LN2=math.log(2.0))

        def m1(fn):
            return self.native(lambda this, a: fn(to_number(a[0]) if a else math.nan))

        def safe(fn):
            def g(x):
                try:
                    return float(fn(x))
                except (ValueError, OverflowError):
                    return math.nan if not (fn is math.exp and x > 0) else math.inf
            return g

        def js_log(x):
            if x != x or x < 0:
                return math.nan
            if x == 0:
                return -math.inf
            return math.log(x) if not math.isinf(x) else math.inf

        def js_sqrt(x):
            return math.nan if (x != x or x < 0) else (math.inf if math.isinf(x) else math.sqrt(x))

        def js_floor(x):
            return x if (x != x or math.isinf(x)) else float(math.floor(x))

        def js_pow(this, a):
            x, y = to_number(a[0]), to_number(a[1])
            with np.errstate(all="ignore"):
                return float(np.power(np.float64(x), np.float64(y)))

        def js_trig(fn):
            return lambda x: math.nan if (x != x or math.isinf(x)) else fn(x)

        M.props["sqrt"] = m1(js_sqrt)
        M.props["cos"] = m1(js_trig(math.cos))
        M.props["sin"] = m1(js_trig(math.sin))
        M.props["atan"] = m1(lambda x: math.atan(x) if x == x else math.nan)
        M.props["abs"] = m1(abs)
        M.props["log"] = m1(js_log)
        M.props["exp"] = m1(safe(math.exp))
        M.props["floor"] = m1(js_floor)
        M.props["ceil"] = m1(lambda x: x if (x != x or math.isinf(x)) else float(math.ceil(x)))
        M.props["round"] = m1(lambda x: x if (x != x or math.isinf(x)) else float(math.floor(x + 0.5)))
        M.props["pow"] = self.native(js_pow)
        M.props["max"] = self.native(lambda this, a: max([to_number(x) for x in a], default=-math.inf))
        M.props["min"] = self.native(lambda this, a: min([to_number(x) for x in a], default=math.inf))
        G["Math"] = M
        G["undefined"] = undefined
        G["NaN"] = math.nan
        G["Infinity"] = math.inf
        console = JSObject(op)
        console.props["log"] = self.native(lambda this, a: print(*[to_str(x) for x in a]) or undefined)
        console.props["error"] = console.props["log"]
        G["console"] = console
        G["isNaN"] = self.native(lambda this, a: to_number(a[0]) != to_number(a[0]))

    def _has_own(self, this, key):
        k = to_key(key)
        if isinstance(this, JSObject):
            return this.has_own(k)
        return False

    def _array_like(self, v):
        if isinstance(v, JSArray):
            return list(v.items)
        if isinstance(v, JSTypedArray):
            return [float(x) for x in v.data]
        return []

    # -- modules (CommonJS, rooted at the reference checkout)
    def require(self, path, base_dir=None):
        p = path if path.endswith(".js") else path + ".js"
        full = os.path.normpath(os.path.join(base_dir if base_dir is not None else self.root, p))
        if full in self.modules:
            return self.modules[full].get("exports")
        module = JSObject(self.object_proto)
        exports = JSObject(self.object_proto)
        module.put("exports", exports)
        self.modules[full] = module
        src = open(full, encoding="utf-8").read()
        for old, new in getattr(self, "source_edits", {}).get(os.path.basename(full), []):
            assert src.count(old) == 1, (full, old)  # (a constant of the module replaced for a parameter study)
            src = src.replace(old, new)
        ast = Parser(src).program()
        env = Env(self.global_env)
        d = os.path.dirname(full)
        env.vars["module"] = module
        env.vars["exports"] = exports
        env.vars["require"] = self.native(lambda this, a: self.require(to_str(a[0]), d), "require")
        self.run_body(ast, env, exports)
        return module.get("exports")

    def run_source(self, src, env=None, this=undefined):
        env = env or self.global_env
        return self.run_body(Parser(src).program(), env, this)

    def run_body(self, ast, env, this):
        names = []
        _hoist(ast, names)
        for n in names:
            env.vars.setdefault(n, undefined)
        for s in ast[1]:
            if s[0] == "funcdecl":
                env.vars[s[1]] = JSFunction(self, s[2][2], s[2][3], env, s[1])
        try:
            self.exec_block(ast[1], env, this)
        except _Return as r:
            return r.value
        return undefined

    def call_function(self, fn, this, args):
        env = Env(fn.env)
        for i, p in enumerate(fn.params):
            env.vars[p] = args[i] if i < len(args) else undefined
        arguments = JSArray(self, args)
        env.vars["arguments"] = arguments
        if fn.name and fn.name not in env.vars:
            env.vars[fn.name] = fn
        return self.run_body(fn.body, env, this)

    # -- statements
    def exec_block(self, stmts, env, this):
        for s in stmts:
            self.exec(s, env, this)

    def exec(self, s, env, this):
        k = s[0]
        if k == "expr":
            self.eval(s[1], env, this)
        elif k == "var":
            for name, init in s[1]:
                if init is not None:
                    env.lookup(name).vars[name] = self.eval(init, env, this)
        elif k == "if":
            if to_boolean(self.eval(s[1], env, this)):
                self.exec(s[2], env, this)
            elif s[3] is not None:
                self.exec(s[3], env, this)
        elif k == "for":
            if s[1] is not None:
                self.exec(s[1], env, this)
            while s[2] is None or to_boolean(self.eval(s[2], env, this)):
                try:
                    self.exec(s[4], env, this)
                except _Break:
                    break
                except _Continue:
                    pass
                if s[3] is not None:
                    self.eval(s[3], env, this)
        elif k == "while":
            while to_boolean(self.eval(s[1], env, this)):
                try:
                    self.exec(s[2], env, this)
                except _Break:
                    break
                except _Continue:
                    pass
        elif k == "dowhile":
            while True:
                try:
                    self.exec(s[2], env, this)
                except _Break:
                    break
                except _Continue:
                    pass
                if not to_boolean(self.eval(s[1], env, this)):
                    break
        elif k == "forin":
            obj = self.eval(s[2], env, this)
            keys = list(obj.props.keys()) if isinstance(obj, JSObject) else []
            if isinstance(obj, JSArray):
                keys = [str(i) for i in range(len(obj.items))] + keys
            for key in keys:
                env.lookup(s[1]).vars[s[1]] = key
                try:
                    self.exec(s[3], env, this)
                except _Break:
                    break
                except _Continue:
                    pass
        elif k == "block":
            self.exec_block(s[1], env, this)
        elif k == "return":
            raise _Return(self.eval(s[1], env, this) if s[1] is not None else undefined)
        elif k == "break":
            raise _Break()
        elif k == "continue":
            raise _Continue()
        elif k == "funcdecl" or k == "empty":
            pass
        elif k == "throw":
            raise JSError(self.eval(s[1], env, this))
        elif k == "try":
            try:
                self.exec(s[1], env, this)
            except JSError as e:
                if s[3] is None:
                    raise
                cenv = Env(env)
                cenv.vars[s[2]] = e.value
                self.exec(s[3], cenv, this)
            finally:
                if s[4] is not None:
                    self.exec(s[4], env, this)
        else:
            raise NotImplementedError(k)

    # -- expressions
    def get_member(self, obj, key):
        if isinstance(obj, JSObject):
            return obj.get(key)
        if isinstance(obj, str):
            if key == "length":
                return float(len(obj))
            i = _array_index(key)
            return obj[i] if i is not None and i < len(obj) else undefined
        if isinstance(obj, (int, float)) and not isinstance(obj, bool):
            return self.number_proto.get(key)
        if obj is undefined or obj is None:
            raise JSError("TypeError: Cannot read property '%s' of %s" % (key, to_str(obj)))
        return undefined

    def put_ref(self, node, value, env, this):
        if node[0] == "id":
            e = env.lookup(node[1])
            (e or self.global_env).vars[node[1]] = value
        elif node[0] == "member":
            obj = self.eval(node[1], env, this)
            key = to_key(self.eval(node[2], env, this))
            if isinstance(obj, JSObject):
                obj.put(key, value)
            elif obj is undefined or obj is None:
                raise JSError("TypeError: Cannot set property '%s' of %s" % (key, to_str(obj)))
        else:
            raise JSError("ReferenceError: invalid assignment target")

    def eval(self, n, env, this):
        k = n[0]
        if k == "num" or k == "str" or k == "bool":
            return n[1]
        if k == "id":
            e = env.lookup(n[1])
            if e is None:
                raise JSError("ReferenceError: %s is not defined" % n[1])
            return e.vars[n[1]]
        if k == "member":
            obj = self.eval(n[1], env, this)
            key = self.eval(n[2], env, this)
            if isinstance(obj, JSTypedArray) and isinstance(key, float):
                i = int(key)
                if i == key and 0 <= i < len(obj.data):
                    return float(obj.data[i])
            return self.get_member(obj, to_key(key))
        if k == "bin":
            return self.binop(n[1], n[2], n[3], env, this)
        if k == "assign":
            op = n[1]
            if op == "=":
                v = self.eval(n[3], env, this)
            else:
                v = self.arith(op[:-1], self.eval(n[2], env, this), self.eval(n[3], env, this))
            self.put_ref(n[2], v, env, this)
            return v
        if k == "call":
            callee = n[1]
            if callee[0] == "member":
                obj = self.eval(callee[1], env, this)
                fn = self.get_member(obj, to_key(self.eval(callee[2], env, this)))
                this_arg = obj
            else:
                fn = self.eval(callee, env, this)
                this_arg = undefined
            args = [self.eval(a, env, this) for a in n[2]]
            if not isinstance(fn, JSFunction):
                raise JSError("TypeError: %s is not a function" % (to_str(fn),))
            return fn.call(this_arg, args)
        if k == "this":
            return this
        if k == "function":
            return JSFunction(self, n[2], n[3], env, n[1])
        if k == "unary":
            op = n[1]
            if op == "typeof":
                if n[2][0] == "id" and env.lookup(n[2][1]) is None:
                    return "undefined"
                return js_typeof(self.eval(n[2], env, this))
            v = self.eval(n[2], env, this)
            if op == "!":
                return not to_boolean(v)
            if op == "-":
                return -to_number(v)
            if op == "+":
                return to_number(v)
            if op == "~":
                return float(~to_int32(v))
            if op == "void":
                return undefined
            raise NotImplementedError(op)
        if k == "update":
            old = to_number(self.eval(n[3], env, this))
            new = old + 1 if n[1] == "++" else old - 1
            self.put_ref(n[3], new, env, this)
            return new if n[2] else old
        if k == "cond":
            return self.eval(n[2] if to_boolean(self.eval(n[1], env, this)) else n[3], env, this)
        if k == "new":
            ctor = self.eval(n[1], env, this)
            args = [self.eval(a, env, this) for a in n[2]]
            if not isinstance(ctor, JSFunction):
                raise JSError("TypeError: not a constructor")
            if ctor.native is not None:
                proto = ctor.props.get("prototype")
                obj = JSObject(proto if isinstance(proto, JSObject) else self.object_proto)
                r = ctor.native(obj, args)
                return r if isinstance(r, JSObject) else obj
            proto = ctor.get("prototype")
            obj = JSObject(proto if isinstance(proto, JSObject) else self.object_proto)
            r = ctor.call(obj, args)
            return r if isinstance(r, JSObject) else obj
        if k == "object":
            o = JSObject(self.object_proto)
            for key, v in n[1]:
                o.props[key] = self.eval(v, env, this)
            return o
        if k == "array":
            return JSArray(self, [self.eval(x, env, this) for x in n[1]])
        if k == "null":
            return None
        if k == "comma":
            self.eval(n[1], env, this)
            return self.eval(n[2], env, this)
        raise NotImplementedError(k)

    def binop(self, op, a, b, env, this):
        if op == "&&":
            l = self.eval(a, env, this)
            return self.eval(b, env, this) if to_boolean(l) else l
        if op == "||":
            l = self.eval(a, env, this)
            return l if to_boolean(l) else self.eval(b, env, this)
        l, r = self.eval(a, env, this), self.eval(b, env, this)
        if op in ("===", "!=="):
            eq = self.strict_equals(l, r)
            return eq if op == "===" else not eq
        if op in ("==", "!="):
            eq = self.loose_equals(l, r)
            return eq if op == "==" else not eq
        if op in ("<", ">", "<=", ">="):
            if isinstance(l, str) and isinstance(r, str):
                return {"<": l < r, ">": l > r, "<=": l <= r, ">=": l >= r}[op]
            x, y = to_number(l), to_number(r)
            if x != x or y != y:
                return False
            return {"<": x < y, ">": x > y, "<=": x <= y, ">=": x >= y}[op]
        if op == "instanceof":
            proto = r.get("prototype") if isinstance(r, JSObject) else None
            o = l.proto if isinstance(l, JSObject) else None
            while o is not None:
                if o is proto:
                    return True
                o = o.proto
            return False
        if op == "in":
            return isinstance(r, JSObject) and (r.get(to_key(l)) is not undefined or r.has_own(to_key(l)))
        return self.arith(op, l, r)

    @staticmethod
    def strict_equals(l, r):
        if isinstance(l, bool) or isinstance(r, bool):
            return isinstance(l, bool) and isinstance(r, bool) and l == r
        if isinstance(l, (int, float)) and isinstance(r, (int, float)):
            return l == r
        if isinstance(l, str) and isinstance(r, str):
            return l == r
        return l is r

    def loose_equals(self, l, r):
        if (l is undefined or l is None) and (r is undefined or r is None):
            return True
        if (l is undefined or l is None) or (r is undefined or r is None):
            return False
        if isinstance(l, JSObject) or isinstance(r, JSObject):
            return l is r
        if isinstance(l, str) and isinstance(r, str):
            return l == r
        return to_number(l) == to_number(r)

    @staticmethod
    def arith(op, l, r):
        if op == "+":
            if isinstance(l, str) or isinstance(r, str) or (isinstance(l, JSObject) or isinstance(r, JSObject)):
                return to_str(l) + to_str(r)
            return to_number(l) + to_number(r)
        if op in ("&", "|", "^"):
            x, y = to_int32(l), to_int32(r)
            return float({"&": x & y, "|": x | y, "^": x ^ y}[op])
        if op == "<<":
            v = (to_int32(l) << (to_uint32(r) & 31)) & 0xFFFFFFFF
            return float(v - 0x100000000 if v >= 0x80000000 else v)
        if op == ">>":
            return float(to_int32(l) >> (to_uint32(r) & 31))
        if op == ">>>":
            return float(to_uint32(l) >> (to_uint32(r) & 31))
        x, y = to_number(l), to_number(r)
        if op == "-":
            return x - y
        if op == "*":
            return x * y
        if op == "/":
            if y == 0:
                if x != x or x == 0:
                    return math.nan
                neg = (math.copysign(1.0, x) < 0) != (math.copysign(1.0, y) < 0)
                return -math.inf if neg else math.inf
            return x / y
        if op == "%":
            if y == 0 or x != x or y != y or math.isinf(x):
                return math.nan
            if math.isinf(y):
                return x
            return math.fmod(x, y)
        raise NotImplementedError(op)

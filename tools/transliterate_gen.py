#!/usr/bin/env python3
"""Mechanical Rust -> C++ transliteration of the reference's machine-generated circuit solvers.

TEST INFRASTRUCTURE.  The melange code generator emits `gen_preamp.rs`, `gen_tremolo.rs` and `gen_power_amp.rs` in a small,
perfectly regular subset of Rust.  This tool parses that subset (a real tokenizer + recursive-descent parser, no per-line
pattern matching) and prints the same program as C++17 (+ GNU statement expressions for Rust's block / if expressions), one
namespace per file, on top of the helper prelude `oracle/rs_prelude.hpp`.  No line of the solvers is restated by hand: operation
order, constants, casts and control flow are whatever the Rust source says.  Compiled with `-ffp-contract=off` against glibc
libm (what Rust's std calls on Linux), the result is used

  * to pin the hand-written oracle (oracle/ow_preamp.hpp, ow_tremolo.hpp) bit for bit on random states
    (tests/test_transliterated_solvers.py), removing human transcription risk from the dominant component, and
  * as the oracle of the melange power amplifier (gen_power_amp.rs, 12 k lines), which has no hand restatement.

Nothing is copied into the repository: the output goes to oracle/_ref/ (git-ignored) and is regenerated from
/root/reference by `make -C oracle ref`.  Usage: transliterate_gen.py <file.rs> <namespace> <out.hpp>
"""
import re
import sys

# ----------------------------------------------------------------------------------------------------------------------
# tokenizer
TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lcomment>//[^\n]*)
  | (?P<bcomment>/\*.*?\*/)
  | (?P<num>0x[0-9A-Fa-f_]+(?:_?(?:u8|u16|u32|u64|usize|i8|i16|i32|i64|isize))?
        | \d[\d_]*(?:\.\d[\d_]*)?(?:[eE][+-]?\d+)?(?:_?(?:f64|f32|u8|u16|u32|u64|usize|i8|i16|i32|i64|isize))?)
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<ident>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op>\.\.=|<<=|>>=|::|->|=>|\.\.|<<|>>|<=|>=|==|!=|&&|\|\||\+=|-=|\*=|/=|%=|\^=|&=|\|=|[-+*/%^!&|=<>@.,;:#$?~(){}\[\]])
""", re.S | re.X)

KEYWORDS = {"let", "mut", "fn", "pub", "const", "static", "struct", "impl", "for", "in", "if", "else", "while", "loop", "return", "break",
            "continue", "as", "use", "mod", "match", "true", "false", "self", "Self", "crate", "super", "where", "type", "enum", "trait", "ref"}


class Tok:
    __slots__ = ("kind", "val", "pos")

    def __init__(self, kind, val, pos):
        self.kind, self.val, self.pos = kind, val, pos

    def __repr__(self):
        return f"{self.kind}:{self.val}"


def tokenize(src):
    toks, i = [], 0
    while i < len(src):
        m = TOKEN_RE.match(src, i)
        if not m:
            raise SyntaxError(f"cannot tokenize at {src[i:i+40]!r}")
        i = m.end()
        k = m.lastgroup
        if k in ("ws", "lcomment", "bcomment"):
            continue
        toks.append(Tok(k, m.group(k), m.start()))
    toks.append(Tok("eof", "", len(src)))
    return toks


# ----------------------------------------------------------------------------------------------------------------------
# parser -> AST (tuples)
class Parser:
    def __init__(self, toks, src):
        self.t, self.i, self.src = toks, 0, src

    def peek(self, k=0):
        return self.t[self.i + k]

    def at(self, val, k=0):
        return self.t[self.i + k].val == val and self.t[self.i + k].kind in ("op", "ident")

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def expect(self, val):
        tok = self.next()
        if tok.val != val:
            line = self.src.count("\n", 0, tok.pos) + 1
            raise SyntaxError(f"line {line}: expected {val!r}, got {tok.val!r}")
        return tok

    def accept(self, val):
        if self.at(val):
            self.i += 1
            return True
        return False

    def line(self):
        return self.src.count("\n", 0, self.peek().pos) + 1

    # ---- items
    def skip_attrs(self):
        """Skips attributes; returns False when a #[cfg(..)] among them disables the next item / statement.  No --cfg flag is
        set in the reference's build, so cfg(NAME) is false and cfg(not(NAME)) is true."""
        enabled = True
        while self.at("#"):
            self.next()
            self.accept("!")
            start = self.i
            self.skip_balanced("[", "]")
            words = [t.val for t in self.t[start + 1:self.i - 1]]
            if words and words[0] == "cfg":
                enabled = enabled and (len(words) > 2 and words[2] == "not")
        return enabled

    def skip_balanced(self, o, c):
        self.expect(o)
        depth = 1
        while depth:
            tok = self.next()
            if tok.kind == "eof":
                raise SyntaxError("unbalanced")
            if tok.kind == "op" and tok.val == o:
                depth += 1
            elif tok.kind == "op" and tok.val == c:
                depth -= 1

    def skip_vis(self):
        if self.accept("pub"):
            if self.at("("):
                self.skip_balanced("(", ")")

    def parse_items(self, until="eof"):
        items = []
        while True:
            self.skip_attrs()
            if (until == "eof" and self.peek().kind == "eof") or (until != "eof" and self.at(until)):
                break
            self.skip_vis()
            tok = self.peek()
            if tok.val == "use":
                while not self.accept(";"):
                    self.next()
            elif tok.val in ("const", "static"):
                self.next()
                self.accept("mut")
                name = self.next().val
                self.expect(":")
                ty = self.parse_type()
                self.expect("=")
                e = self.parse_expr()
                self.expect(";")
                items.append(("const", name, ty, e))
            elif tok.val == "struct":
                self.next()
                name = self.next().val
                fields = []
                self.expect("{")
                while not self.accept("}"):
                    self.skip_attrs()
                    self.skip_vis()
                    fname = self.next().val
                    self.expect(":")
                    fields.append((fname, self.parse_type()))
                    self.accept(",")
                items.append(("struct", name, fields))
            elif tok.val == "impl":
                self.next()
                a = self.next().val
                trait = None
                if self.accept("for"):
                    trait, a = a, self.next().val
                self.expect("{")
                fns = self.parse_items(until="}")
                self.expect("}")
                items.append(("impl", a, trait, fns))
            elif tok.val == "fn":
                items.append(self.parse_fn())
            elif tok.val == "mod":
                self.next()
                self.next()
                if not self.accept(";"):
                    self.skip_balanced("{", "}")
            else:
                raise SyntaxError(f"line {self.line()}: unexpected item token {tok.val!r}")
        return items

    def parse_fn(self):
        self.expect("fn")
        name = self.next().val
        generic = False
        if self.at("<"):
            generic = True
            depth = 0
            while True:
                tok = self.next()
                if tok.val == "<":
                    depth += 1
                elif tok.val == ">":
                    depth -= 1
                    if depth == 0:
                        break
        self.expect("(")
        params = []
        self_kind = None
        while not self.accept(")"):
            if self.at("&") and (self.at("self", 1) or (self.at("mut", 1) and self.at("self", 2))):
                self.next()
                self_kind = "mut" if self.accept("mut") else "ref"
                self.expect("self")
            elif self.at("self"):
                self.next()
                self_kind = "val"
            else:
                self.accept("mut")
                pname = self.next().val
                self.expect(":")
                params.append((pname, self.parse_type()))
            self.accept(",")
        ret = None
        if self.accept("->"):
            ret = self.parse_type()
        start = self.i
        if generic:
            self.skip_balanced("{", "}")
            return ("fn", name, params, ret, None, self_kind, True)
        try:
            body = self.parse_block()
        except Unsupported:
            self.i = start
            self.skip_balanced("{", "}")
            body = None
        return ("fn", name, params, ret, body, self_kind, False)

    # ---- types
    def parse_type(self):
        if self.accept("&"):
            m = self.accept("mut")
            return ("ref", self.parse_type(), m)
        if self.accept("["):
            el = self.parse_type()
            self.expect(";")
            n = self.parse_expr()
            self.expect("]")
            return ("array", el, n)
        if self.accept("("):
            parts = []
            while not self.accept(")"):
                parts.append(self.parse_type())
                self.accept(",")
            return ("tuple", parts)
        name = self.next().val
        while self.accept("::"):
            name += "::" + self.next().val
        args = []
        if self.at("<"):
            self.next()
            while not self.accept(">"):
                args.append(self.parse_type())
                self.accept(",")
        return ("path", name, args)

    # ---- blocks and statements
    def parse_block(self):
        self.expect("{")
        stmts, tail = [], None
        while not self.at("}"):
            if not self.skip_attrs():  # #[cfg(..)] that is off in the reference's build: parse and drop the statement
                if self.at("let"):
                    while not self.accept(";"):
                        self.next()
                else:
                    self.parse_expr(stmt=True)
                    self.accept(";")
                continue
            if self.accept(";"):
                continue
            if self.at("let"):
                self.next()
                pat = self.parse_pattern()
                ty = None
                if self.accept(":"):
                    ty = self.parse_type()
                init = None
                if self.accept("="):
                    init = self.parse_expr()
                self.expect(";")
                stmts.append(("let", pat, ty, init))
                continue
            if self.at("const"):
                self.next()
                name = self.next().val
                self.expect(":")
                ty = self.parse_type()
                self.expect("=")
                e = self.parse_expr()
                self.expect(";")
                stmts.append(("localconst", name, ty, e))
                continue
            if self.peek().val in ("if", "for", "while", "loop", "{") and self.peek().kind in ("ident", "op"):
                # block-like expression statement: ends at its closing brace (no postfix / binary continuation)
                e = self.parse_primary(False)
                if self.at("}") and e[0] not in ("for", "while", "loop") and not (e[0] == "if" and e[3] is None):
                    tail = e
                else:
                    self.accept(";")
                    stmts.append(("expr", e))
                continue
            e = self.parse_expr(stmt=True)
            if self.accept(";"):
                stmts.append(("expr", e))
            elif self.at("}") and e[0] not in ("for", "while", "loop") and not (e[0] == "if" and e[3] is None):
                tail = e
            elif e[0] in ("if", "for", "while", "loop", "block", "iflet"):
                stmts.append(("expr", e))
            else:
                raise SyntaxError(f"line {self.line()}: expected ';' after expression, got {self.peek().val!r}")
        self.expect("}")
        return ("block", stmts, tail)

    def parse_pattern(self):
        if self.accept("("):
            parts = []
            while not self.accept(")"):
                parts.append(self.parse_pattern())
                self.accept(",")
            return ("ptuple", parts)
        self.accept("&")
        m = self.accept("mut")
        return ("pident", self.next().val, m)

    # ---- expressions (precedence climbing)
    BINOPS = [("||",), ("&&",), ("==", "!=", "<", ">", "<=", ">="), ("|",), ("^",), ("&",), ("<<", ">>"), ("+", "-"), ("*", "/", "%")]
    ASSIGN = ("=", "+=", "-=", "*=", "/=", "%=", "^=", "&=", "|=", "<<=", ">>=")

    def parse_expr(self, stmt=False, nostruct=False):
        lhs = self.parse_range(nostruct)
        if self.peek().kind == "op" and self.peek().val in self.ASSIGN:
            op = self.next().val
            rhs = self.parse_expr(nostruct=nostruct)
            return ("assign", op, lhs, rhs)
        return lhs

    def parse_range(self, nostruct):
        lhs = self.parse_bin(0, nostruct)
        if self.at("..") or self.at("..="):
            incl = self.next().val == "..="
            rhs = self.parse_bin(0, nostruct)
            return ("range", lhs, rhs, incl)
        return lhs

    def parse_bin(self, level, nostruct):
        if level == len(self.BINOPS):
            return self.parse_cast(nostruct)
        lhs = self.parse_bin(level + 1, nostruct)
        while self.peek().kind == "op" and self.peek().val in self.BINOPS[level]:
            # `a < b` vs generic `<`: generics never follow a value expression in this subset
            op = self.next().val
            rhs = self.parse_bin(level + 1, nostruct)
            lhs = ("bin", op, lhs, rhs)
        return lhs

    def parse_cast(self, nostruct):
        e = self.parse_unary(nostruct)
        while self.at("as"):
            self.next()
            e = ("cast", e, self.parse_type())
        return e

    def parse_unary(self, nostruct):
        if self.at("-"):
            self.next()
            return ("un", "-", self.parse_unary(nostruct))
        if self.at("!"):
            self.next()
            return ("un", "!", self.parse_unary(nostruct))
        if self.at("*"):
            self.next()
            return ("deref", self.parse_unary(nostruct))
        if self.at("&"):
            self.next()
            self.accept("mut")
            return ("addr", self.parse_unary(nostruct))
        if self.at("&&"):
            self.next()
            self.accept("mut")
            return ("addr", self.parse_unary(nostruct))
        return self.parse_postfix(self.parse_primary(nostruct), nostruct)

    def parse_args(self):
        self.expect("(")
        args = []
        while not self.accept(")"):
            args.append(self.parse_expr())
            self.accept(",")
        return args

    def parse_postfix(self, e, nostruct):
        while True:
            if self.at("."):
                nxt = self.peek(1)
                if nxt.kind == "num":
                    self.next()
                    e = ("tupidx", e, int(self.next().val))
                    continue
                self.next()
                name = self.next().val
                if self.at("::"):  # turbofish on a method: not in this subset
                    raise Unsupported("method turbofish")
                if self.at("("):
                    e = ("method", e, name, self.parse_args())
                else:
                    e = ("field", e, name)
            elif self.at("["):
                self.next()
                idx = self.parse_expr()
                self.expect("]")
                e = ("index", e, idx)
            elif self.at("("):
                e = ("call", e, self.parse_args())
            elif self.at("?"):
                raise Unsupported("? operator")
            else:
                return e

    def parse_primary(self, nostruct):
        tok = self.peek()
        if tok.kind == "num":
            self.next()
            return ("num", tok.val)
        if tok.kind == "str":
            raise Unsupported("string literal")
        if tok.val == "(":
            self.next()
            if self.accept(")"):
                return ("tuple", [])
            first = self.parse_expr()
            if self.accept(")"):
                return ("paren", first)
            parts = [first]
            while self.accept(","):
                if self.at(")"):
                    break
                parts.append(self.parse_expr())
            self.expect(")")
            return ("tuple", parts)
        if tok.val == "[":
            self.next()
            if self.accept("]"):
                return ("array", [])
            first = self.parse_expr()
            if self.accept(";"):
                n = self.parse_expr()
                self.expect("]")
                return ("repeat", first, n)
            parts = [first]
            while self.accept(","):
                if self.at("]"):
                    break
                parts.append(self.parse_expr())
            self.expect("]")
            return ("array", parts)
        if tok.val == "{":
            return self.parse_block()
        if tok.val == "if":
            self.next()
            if self.at("let"):
                self.next()
                # if let Some(x) = expr { } else { }
                ctor = self.next().val
                self.expect("(")
                binder = self.parse_pattern()
                self.expect(")")
                self.expect("=")
                scrut = self.parse_expr(nostruct=True)
                then = self.parse_block()
                els = None
                if self.accept("else"):
                    els = self.parse_block() if self.at("{") else ("block", [], self.parse_primary(False))
                return ("iflet", ctor, binder, scrut, then, els)
            cond = self.parse_expr(nostruct=True)
            then = self.parse_block()
            els = None
            if self.accept("else"):
                els = self.parse_block() if self.at("{") else ("block", [], self.parse_primary(False))
            return ("if", cond, then, els)
        if tok.val == "for":
            self.next()
            pat = self.parse_pattern()
            self.expect("in")
            it = self.parse_expr(nostruct=True)
            return ("for", pat, it, self.parse_block())
        if tok.val == "while":
            self.next()
            cond = self.parse_expr(nostruct=True)
            return ("while", cond, self.parse_block())
        if tok.val == "loop":
            self.next()
            return ("loop", self.parse_block())
        if tok.val == "return":
            self.next()
            if self.at(";") or self.at("}"):
                return ("return", None)
            return ("return", self.parse_expr())
        if tok.val == "break":
            self.next()
            return ("break",)
        if tok.val == "continue":
            self.next()
            return ("continue",)
        if tok.val == "|" or tok.val == "||":
            # closure |pat, pat| expr
            params = []
            if self.next().val == "|":
                while not self.accept("|"):
                    byref = not self.at("&")
                    params.append((self.parse_pattern(), byref))
                    self.accept(",")
            return ("closure", params, self.parse_expr())
        if tok.val == "match":
            raise Unsupported("match")
        if tok.kind == "ident":
            self.next()
            path = [tok.val]
            generics = None
            while self.at("::"):
                self.next()
                if self.at("<"):
                    self.next()
                    generics = []
                    while not self.accept(">"):
                        generics.append(self.parse_type())
                        self.accept(",")
                else:
                    path.append(self.next().val)
            if self.at("!"):
                nxt = self.peek(1)
                if nxt.val in ("(", "[", "{") and self.peek().pos + 1 == nxt.pos:
                    raise Unsupported("macro " + tok.val)
            if self.at("{") and not nostruct and (path[-1][0].isupper() and not path[-1].isupper()):
                self.next()
                fields = []
                while not self.accept("}"):
                    fname = self.next().val
                    if self.accept(":"):
                        fields.append((fname, self.parse_expr()))
                    else:
                        fields.append((fname, ("path", [fname], None)))
                    self.accept(",")
                return ("structlit", path, fields)
            return ("path", path, generics)
        raise SyntaxError(f"line {self.line()}: unexpected token {tok.val!r}")


class Unsupported(Exception):
    pass


# ----------------------------------------------------------------------------------------------------------------------
# C++ emitter
CPP_RESERVED = {"default", "new", "delete", "this", "class", "template", "register", "auto", "signed", "unsigned", "switch", "case", "double", "float",
                "int", "long", "short", "char", "union", "goto", "operator", "private", "public", "protected", "friend", "typename", "namespace",
                "and", "or", "not", "xor", "inline", "extern", "volatile", "asm", "bool", "do", "try", "catch", "throw", "using", "virtual", "explicit",
                "export", "typedef", "sizeof", "void", "static_cast", "signal", "div", "exp", "log", "abs", "y0", "y1", "j0", "j1", "jn", "yn", "gamma",
                "index", "time", "remainder", "free", "malloc", "exit", "rand", "system", "INFINITY", "NAN", "EOF", "NULL", "M_PI", "M_E", "HUGE_VAL"}
PRIM = {"f64": "double", "f32": "float", "usize": "size_t", "isize": "ptrdiff_t", "u64": "uint64_t", "u32": "uint32_t", "u16": "uint16_t", "u8": "uint8_t",
        "i64": "int64_t", "i32": "int32_t", "i16": "int16_t", "i8": "int8_t", "bool": "bool"}


def cname(n):
    return n + "_" if n in CPP_RESERVED else n


class Emitter:
    def __init__(self, items):
        self.items = items
        self.structs = {it[1]: it for it in items if it[0] == "struct"}
        self.methods = {}   # struct -> set of method names
        for it in items:
            if it[0] == "impl":
                self.methods.setdefault(it[1], set()).update(f[1] for f in it[3] if f[0] == "fn")
        self.all_methods = set().union(*self.methods.values()) if self.methods else set()
        self.out = []
        self.scopes = []
        self.used_names = set()
        self.tmp = 0
        self.cur_struct = None
        self.ptr_vars = set()
        self.skipped = []

    # ---- names / scopes
    def push(self):
        self.scopes.append({})

    def pop(self):
        self.scopes.pop()

    def declare(self, name, pointer=False):
        c = cname(name)
        if c in self.used_names:
            k = 1
            while f"{c}__{k}" in self.used_names:
                k += 1
            c = f"{c}__{k}"
        self.used_names.add(c)
        self.scopes[-1][name] = c
        if pointer:
            self.ptr_vars.add(c)
        return c

    def lookup(self, name):
        for sc in reversed(self.scopes):
            if name in sc:
                return sc[name]
        return cname(name)

    def fresh(self, base="__t"):
        self.tmp += 1
        return f"{base}{self.tmp}"

    # ---- types
    def ty(self, t, param=False):
        k = t[0]
        if k == "ref":
            inner = self.ty(t[1])
            return (inner + "&") if t[2] else ("const " + inner + "&")
        if k == "array":
            return f"std::array<{self.ty(t[1])}, {self.expr(t[2])}>"
        if k == "tuple":
            return "std::tuple<" + ", ".join(self.ty(x) for x in t[1]) + ">"
        name, args = t[1], t[2]
        if name in PRIM:
            return PRIM[name]
        if name == "Self":
            return self.cur_struct
        if name == "Option":
            return f"rs::Option<{self.ty(args[0])}>"
        return cname(name.split("::")[-1])

    # ---- expressions
    def num(self, v):
        m = re.match(r"^(.*?)(?:_?(f64|f32|u8|u16|u32|u64|usize|i8|i16|i32|i64|isize))?$", v)
        body, suf = m.group(1).replace("_", ""), m.group(2)
        if body.startswith("0x"):
            val = int(body, 16)
            lit = f"{body}ull" if val > 0x7FFFFFFF else body
            return f"(({PRIM[suf]}){lit})" if suf else lit
        is_float = any(c in body for c in ".eE")
        if suf in ("f64", "f32"):
            if not is_float:
                body += ".0"
            return body if suf == "f64" else body + "f"
        if is_float:
            return body
        val = int(body)
        lit = f"{body}ull" if val > 0x7FFFFFFF else body
        return f"(({PRIM[suf]}){lit})" if suf else lit

    def expr(self, e, want=True):
        k = e[0]
        if k == "num":
            return self.num(e[1])
        if k == "paren":
            return "(" + self.expr(e[1]) + ")"
        if k == "path":
            path, generics = e[1], e[2]
            if len(path) == 1:
                n = path[0]
                if n in ("true", "false"):
                    return n
                if n == "None":
                    return "rs::none"
                if n == "self":
                    return "self"
                s = self.lookup(n)
            else:
                head = path[0]
                if head in PRIM:
                    s = f"rs::{head}_::{cname(path[-1])}"
                elif head in ("std", "core"):
                    s = "rs::" + "_".join(path[1:])
                elif head == "Self":
                    s = f"{self.cur_struct}::{cname(path[-1])}"
                else:
                    s = "::".join(cname(p) for p in path)
            if generics:
                s += "<" + ", ".join(self.ty(g) for g in generics) + ">"
            return s
        if k == "bin":
            return f"({self.expr(e[2])} {e[1]} {self.expr(e[3])})"
        if k == "un":
            return f"({e[1]}{self.expr(e[2])})"
        if k == "deref":
            inner = self.expr(e[1])
            return f"(*{inner})" if inner in self.ptr_vars else inner
        if k == "addr":
            return self.expr(e[1])
        if k == "cast":
            return f"rs::as_<{self.ty(e[2])}>({self.expr(e[1])})"
        if k == "field":
            return f"{self.expr(e[1])}.{cname(e[2])}"
        if k == "tupidx":
            return f"std::get<{e[2]}>({self.expr(e[1])})"
        if k == "index":
            return f"{self.expr(e[1])}[{self.expr(e[2])}]"
        if k == "call":
            f = e[1]
            if f[0] == "path" and f[1] == ["Some"]:
                return f"rs::some({self.expr(e[2][0])})"
            return f"{self.expr(f)}({', '.join(self.expr(a) for a in e[2])})"
        if k == "method":
            recv, name, args = e[1], e[2], e[3]
            if name in self.all_methods and name not in ("abs", "max", "min", "clamp"):
                return f"{self.expr(recv)}.{cname(name)}({', '.join(self.expr(a) for a in args)})"
            if name in ("iter", "iter_mut", "into_iter") and not args:
                return f"rs::iter({self.expr(recv)})"
            if name == "rev" and recv[0] == "paren" and recv[1][0] == "range":
                return ("revrange", recv[1])  # only meaningful as a for-loop iterator
            return f"rs::m_{name}({', '.join([self.expr(recv)] + [self.expr(a) for a in args])})"
        if k == "tuple":
            return "std::make_tuple(" + ", ".join(self.expr(x) for x in e[1]) + ")"
        if k == "array":
            return "rs::arr(" + ", ".join(self.expr(x) for x in e[1]) + ")"
        if k == "repeat":
            v = e[1]
            if v[0] == "path" and v[1] == ["None"]:
                return "rs::Default{}"
            return f"rs::fill<{self.expr(e[2])}>({self.expr(v)})"
        if k == "structlit":
            name = self.cur_struct if e[1] == ["Self"] else cname(e[1][-1])
            t = self.fresh("__s")
            body = "".join(f"{t}.{cname(fn)} = {self.expr(fe)}; " for fn, fe in e[2])
            return f"({{ {name} {t}; {body}{t}; }})"
        if k == "closure":
            self.push()
            ps = []
            for pat, byref in e[1]:
                c = self.declare(pat[1])
                ps.append(f"const auto& {c}")
            body = self.expr(e[2])
            self.pop()
            return f"[&]({', '.join(ps)}) {{ return {body}; }}"
        if k == "assign":
            return f"{self.expr(e[2])} {e[1]} {self.expr(e[3])}"
        if k == "block":
            return self.block_expr(e)
        if k == "if":
            if e[3] is None:
                raise Unsupported("value of if without else")
            return f"(({self.expr(e[1])}) ? {self.block_expr(e[2])} : {self.block_expr(e[3])})"
        if k == "range":
            raise Unsupported("range value")
        if k == "return":
            return "({ " + self.stmt_str(("expr", e)) + " 0; })"
        raise Unsupported(f"expression kind {k}: {str(e)[:300]}")

    def expr_typed(self, e, ty):
        """Array literals under a declared type take the declared element type (integer tables would otherwise be `int`)."""
        if ty is not None and ty[0] == "array":
            el = ty[1]
            if e[0] == "array":
                if el[0] == "array":
                    return "rs::arr(" + ", ".join(self.expr_typed(x, el) for x in e[1]) + ")"
                return f"rs::arr_of<{self.ty(el)}>(" + ", ".join(self.expr(x) for x in e[1]) + ")"
            if e[0] == "repeat" and not (e[1][0] == "path" and e[1][1] == ["None"]):
                inner = self.expr_typed(e[1], el) if el[0] == "array" else f"({self.ty(el)})({self.expr(e[1])})"
                return f"rs::fill<{self.expr(e[2])}>({inner})"
        return self.expr(e)

    def block_expr(self, b):
        """Rust block as a value: GNU statement expression."""
        if b[0] != "block":
            return self.expr(b)
        if not b[1] and b[2] is not None and b[2][0] not in ("if", "block"):
            return "(" + self.expr(b[2]) + ")"
        saved = self.out
        self.out = []
        self.push()
        for s in b[1]:
            self.stmt(s, 0)
        tail = self.expr(b[2]) if b[2] is not None else "0"
        self.pop()
        inner = " ".join(x.strip() for x in self.out)
        self.out = saved
        return f"({{ {inner} {tail}; }})"

    # ---- statements
    def stmt_str(self, s):
        saved = self.out
        self.out = []
        self.stmt(s, 0)
        r = " ".join(x.strip() for x in self.out)
        self.out = saved
        return r

    def w(self, ind, text):
        self.out.append("    " * ind + text)

    def bind_pattern(self, pat, value, ind):
        if pat[0] == "pident":
            c = self.declare(pat[1])
            self.w(ind, f"auto {c} = {value};")
        else:
            t = self.fresh()
            self.w(ind, f"auto {t} = {value};")
            for idx, p in enumerate(pat[1]):
                self.bind_pattern(p, f"std::get<{idx}>({t})", ind)

    def stmt(self, s, ind):
        k = s[0]
        if k == "let":
            pat, ty, init = s[1], s[2], s[3]
            if init is None:
                c = self.declare(pat[1])
                self.w(ind, f"{self.ty(ty)} {c};")
                return
            val = self.expr_typed(init, ty)   # evaluated BEFORE the new binding shadows an old one
            if pat[0] == "pident" and ty is not None:
                c = self.declare(pat[1])
                self.w(ind, f"{self.ty(ty)} {c} = {val};")
            else:
                self.bind_pattern(pat, val, ind)
            return
        if k == "localconst":
            val = self.expr_typed(s[3], s[2])
            c = self.declare(s[1])
            self.w(ind, f"const {self.ty(s[2])} {c} = {val};")
            return
        e = s[1]
        ek = e[0]
        if ek == "if":
            self.w(ind, f"if ({self.expr(e[1])}) {{")
            self.block_stmts(e[2], ind + 1)
            if e[3] is not None:
                if e[3][0] == "block" and not e[3][1] and e[3][2] is not None and e[3][2][0] == "if":
                    self.w(ind, "} else {")
                    self.stmt(("expr", e[3][2]), ind + 1)
                else:
                    self.w(ind, "} else {")
                    self.block_stmts(e[3], ind + 1)
            self.w(ind, "}")
        elif ek == "iflet":
            ctor, binder, scrut, then, els = e[1:]
            t = self.fresh("__o")
            self.w(ind, "{")
            self.w(ind + 1, f"auto {t} = {self.expr(scrut)};")
            self.w(ind + 1, f"if ({t}.is_some) {{")
            self.push()
            c = self.declare(binder[1])
            self.w(ind + 2, f"auto {c} = {t}.value;")
            self.block_stmts(then, ind + 2, scoped=False)
            self.pop()
            if els is not None:
                self.w(ind + 1, "} else {")
                self.block_stmts(els, ind + 2)
            self.w(ind + 1, "}")
            self.w(ind, "}")
        elif ek == "for":
            pat, it, body = e[1], e[2], e[3]
            self.push()
            rev = False
            if it[0] == "method" and it[2] == "rev" and it[1][0] == "paren" and it[1][1][0] == "range":
                it, rev = it[1][1], True
            if it[0] == "range":
                lo, hi = self.expr(it[1]), self.expr(it[2])
                if it[3]:
                    hi = f"(({hi}) + 1)"
                c = self.declare(pat[1])
                if rev:
                    self.w(ind, f"for (size_t {c} = {hi}; {c}-- > (size_t)({lo});) {{")
                else:
                    self.w(ind, f"for (size_t {c} = {lo}; {c} < (size_t)({hi}); {c}++) {{")
            elif it[0] == "method" and it[2] == "iter_mut":
                c = self.declare(pat[1], pointer=True)
                t = self.fresh("__e")
                self.w(ind, f"for (auto& {t} : {self.expr(it[1])}) {{")
                self.w(ind + 1, f"auto* {c} = &{t};")
            else:
                raise Unsupported("for over " + it[0])
            self.block_stmts(body, ind + 1)
            self.w(ind, "}")
            self.pop()
        elif ek == "while":
            self.w(ind, f"while ({self.expr(e[1])}) {{")
            self.block_stmts(e[2], ind + 1)
            self.w(ind, "}")
        elif ek == "loop":
            self.w(ind, "for (;;) {")
            self.block_stmts(e[1], ind + 1)
            self.w(ind, "}")
        elif ek == "block":
            self.w(ind, "{")
            self.block_stmts(e, ind + 1)
            self.w(ind, "}")
        elif ek == "return":
            self.w(ind, "return;" if e[1] is None else f"return {self.expr(e[1])};")
        elif ek == "break":
            self.w(ind, "break;")
        elif ek == "continue":
            self.w(ind, "continue;")
        else:
            self.w(ind, self.expr(e) + ";")

    def block_stmts(self, b, ind, scoped=True):
        """A block in statement position (its value, if any, is discarded)."""
        if scoped:
            self.push()
        for s in b[1]:
            self.stmt(s, ind)
        if b[2] is not None:
            self.stmt(("expr", b[2]), ind)
        if scoped:
            self.pop()

    # ---- items
    def fn(self, f, ind, struct=None):
        _, name, params, ret, body, self_kind, generic = f
        if generic or body is None:
            self.skipped.append(name)
            return
        self.scopes = [{}]
        self.used_names = set()
        self.ptr_vars = set()
        self.cur_struct = struct
        ps = []
        for pn, pt in params:
            c = self.declare(pn)
            ps.append(f"{self.ty(pt)} {c}")
        rt = self.ty(ret) if ret is not None else "void"
        if ret is not None and ret[0] == "ref":
            rt = self.ty(ret)
        static = "static " if (struct and self_kind is None) else ("inline " if not struct else "")
        saved = self.out
        self.out = []
        try:
            self.w(ind, f"{static}{rt} {cname(name)}({', '.join(ps)}) {{")
            if struct and self_kind is not None:
                self.w(ind + 1, "auto& self = *this;")
            self.push()
            for s in body[1]:
                self.stmt(s, ind + 1)
            if body[2] is not None:
                if ret is not None:
                    self.w(ind + 1, f"return {self.expr(body[2])};")
                else:
                    self.stmt(("expr", body[2]), ind + 1)
            self.pop()
            self.w(ind, "}")
            saved.extend(self.out)
        except Unsupported as ex:
            self.skipped.append(f"{name} ({ex})")
        self.out = saved

    def emit(self, ns):
        self.w(0, f"namespace {ns} {{")
        # forward declarations: structs first (methods need the free functions, free functions need the structs)
        consts = [it for it in self.items if it[0] == "const"]
        self.scopes = [{}]
        self.used_names = set()
        for it in consts:
            _, name, ty, e = it
            try:
                if ty[0] == "path" and ty[1] in PRIM:
                    self.w(0, f"static constexpr {self.ty(ty)} {cname(name)} = {self.expr(e)};")
                else:
                    self.w(0, f"static const {self.ty(ty)} {cname(name)} = {self.expr_typed(e, ty)};")
            except Unsupported as ex:
                self.skipped.append(f"const {name} ({ex})")
        for it in self.items:
            if it[0] == "struct":
                self.w(0, f"struct {cname(it[1])};")
        free_fns = [it for it in self.items if it[0] == "fn"]
        STUBS = {"gaussian": "inline double gaussian(Xoshiro256pp&, rs::Option<double>&) { return 0.0; }  // noise is off in every parity run",
                 "seed_noise_rngs": "template <size_t N_> std::array<Xoshiro256pp, N_> seed_noise_rngs(uint64_t) { return {}; }",
                 "seed_noise_rngs_salted": "template <size_t N_> std::array<Xoshiro256pp, N_> seed_noise_rngs_salted(uint64_t, uint64_t) { return {}; }"}
        self.stub_lines = [STUBS[f[1]] for f in free_fns if f[1] in STUBS]
        # struct definitions with method declarations
        for it in self.items:
            if it[0] != "struct":
                continue
            sname = it[1]
            self.cur_struct = cname(sname)
            self.w(0, f"struct {cname(sname)} {{")
            for fname, fty in it[2]:
                self.w(1, f"{self.ty(fty)} {cname(fname)}{{}};")
            for imp in self.items:
                if imp[0] == "impl" and imp[1] == sname:
                    for f in imp[3]:
                        if f[0] == "fn" and not f[6] and f[4] is not None:
                            _, name, params, ret, body, self_kind, generic = f
                            self.scopes = [{}]
                            self.used_names = set()
                            ps = ", ".join(f"{self.ty(pt)} {cname(pn)}" for pn, pt in params)
                            rt = self.ty(ret) if ret is not None else "void"
                            st = "static " if self_kind is None else ""
                            self.w(1, f"{st}{rt} {cname(name)}({ps});")
            self.w(0, "};")
        for l in self.stub_lines:
            self.w(0, l)
        # free function prototypes
        for f in free_fns:
            _, name, params, ret, body, self_kind, generic = f
            if generic or body is None or name in STUBS:
                continue
            self.scopes = [{}]
            self.used_names = set()
            self.cur_struct = None
            ps = ", ".join(f"{self.ty(pt)} {cname(pn)}" for pn, pt in params)
            rt = self.ty(ret) if ret is not None else "void"
            self.w(0, f"inline {rt} {cname(name)}({ps});")
        # bodies
        for f in free_fns:
            if f[1] not in STUBS:
                self.fn(f, 0)
        for imp in self.items:
            if imp[0] == "impl":
                for f in imp[3]:
                    if f[0] != "fn":
                        continue
                    before = len(self.out)
                    self.fn(f, 0, struct=cname(imp[1]))
                    # turn the in-class style header into an out-of-class definition
                    if len(self.out) > before:
                        hdr = self.out[before]
                        hdr = hdr.replace("static ", "", 1)
                        m = re.match(r"^(\s*)(.*?)\s(\w+)\((.*)\) \{$", hdr)
                        self.out[before] = f"{m.group(1)}inline {m.group(2)} {cname(imp[1])}::{m.group(3)}({m.group(4)}) {{"
        self.w(0, f"}}  // namespace {ns}")
        self.out = [l for l in self.out if not any(re.match(rf"^inline .*\b{re.escape(n.split()[0])}\(.*\);$", l) for n in self.skipped)]
        return "\n".join(self.out) + "\n"


def main():
    src_path, ns, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
    src = open(src_path).read()
    toks = tokenize(src)
    items = Parser(toks, src).parse_items()
    em = Emitter(items)
    text = em.emit(ns)
    with open(out_path, "w") as f:
        f.write(f"// GENERATED by tools/transliterate_gen.py from {src_path} -- do not edit, do not commit.\n")
        f.write("// TEST INFRASTRUCTURE: mechanical Rust -> C++ transliteration of the reference's generated solver.\n")
        f.write('#pragma once\n#include "rs_prelude.hpp"\n')
        f.write(text)
    if em.skipped:
        sys.stderr.write(f"{ns}: not transliterated (outside the subset / noise-only): {', '.join(em.skipped)}\n")


if __name__ == "__main__":
    main()

"""A small Erlang evaluator that runs the reference's SOURCE TEXT (raytracer.erl) directly.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): nothing in eraytracer_b200/ may import it.

Why it exists: Erlang/OTP is not installed here, so `oracle.c` / `pyoracle.py` are human restatements
of raytracer.erl:180-614 and every parity claim rested on that reading.  This module takes the
reading out of the loop: it tokenises and parses the module text, and evaluates its functions with
Erlang's semantics —

  * numbers: integers are Python ints (unbounded, like BEAM bignums), floats are IEEE doubles;
    `+ - *` stay integer on two integers and go through double otherwise, `/` is always double,
    `==`/`<`/... compare numerically across int/float, `=:=` and PATTERNS compare exactly
    (the integer 0 does not match 0.0);
  * term order for mixed comparisons: number < atom < fun < tuple < list (erl:319 compares the atom
    `infinity` with a float, and `or` evaluates both operands);
  * records are tagged tuples laid out by the module's own -record declarations (erl:72-81);
  * `math:sqrt/pow/tan/pi` are the platform libm through Python's math module — what BEAM's BIFs call.

What is human-made here is an interpreter of the LANGUAGE, not of the ray tracer, and it is checked
by the reference's own test-suite: `run_tests/0` (erl:736-1133) is executed by this evaluator and
must return `ok` (tests/test_erl_reference.py).  The images it renders from the untouched source are
the golden vectors of tests/golden/erl_reference.json.

Supported subset: everything raytracer.erl uses outside processes and distribution (no `receive`
evaluation, no spawn/pool; those forms are parsed, not run), plus what erl/raytracer_gpu_scenes.erl needs:
based integer literals (16#FF), band/bor/bxor/bsl/bsr on unbounded integers, min/max, and whole-byte
float / unsigned-integer binary segments (`<<F:32/float>> = <<X:32/float>>` rounds to binary32 as BEAM does);
and what erl/raytracer_gpu.erl needs: `$c` literals, `float-native` / `Rest/binary` segments, binary generators in
list comprehensions, exit/1, is_list/1, integer_to_list/1, proplists:get_value, io_lib:format, file:write_file, and
`Module.externals` — Python stand-ins for NIFs — and a SEQUENTIAL process model for render_binary/5: spawn/spawn_link
run the fun to completion at the spawn, `!` appends to a mailbox, `receive` takes the first matching message and fails
if it would have to wait (a valid schedule for programs whose children never wait for their parent; the reference's own
master/worker drivers are still outside the evaluator).
"""
import math
import re
import struct
import sys


# --------------------------------------------------------------------------- terms
class Atom(str):
    """An Erlang atom.  Interned by value: Atom('ok') == Atom('ok'), and never equal to a plain str."""
    __slots__ = ()

    def __repr__(self):
        return str.__str__(self)

    def __eq__(self, other):
        return isinstance(other, Atom) and str.__eq__(self, other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = str.__hash__


class ErlString(list):
    """A string literal: a list of code points (kept distinguishable for io:format)."""

    def text(self):
        return "".join(chr(c) for c in self)


TRUE, FALSE = Atom("true"), Atom("false")


class ErlError(Exception):
    pass


class ErlExit(ErlError):
    """exit(Reason)"""

    def __init__(self, reason):
        ErlError.__init__(self, "exit: %r" % (reason,))
        self.reason = reason


def _type_rank(v):
    if isinstance(v, bool):
        raise ErlError("python bool leaked into Erlang terms")
    if isinstance(v, (int, float)):
        return 0
    if isinstance(v, Atom):
        return 1
    if callable(v):
        return 2
    if isinstance(v, tuple):
        return 3
    if isinstance(v, list):
        return 4
    if isinstance(v, bytes):
        return 5
    raise ErlError("no term order for %r" % (v,))


def term_cmp(a, b):
    """Erlang term order: -1, 0, 1 (numbers compare by value across int/float)."""
    ra, rb = _type_rank(a), _type_rank(b)
    if ra != rb:
        return -1 if ra < rb else 1
    if ra == 0:
        return -1 if a < b else (1 if a > b else 0)
    if ra == 1:
        return -1 if str(a) < str(b) else (1 if str(a) > str(b) else 0)
    if ra == 3:
        if len(a) != len(b):
            return -1 if len(a) < len(b) else 1
        for x, y in zip(a, b):
            c = term_cmp(x, y)
            if c:
                return c
        return 0
    if ra == 4:
        for x, y in zip(a, b):
            c = term_cmp(x, y)
            if c:
                return c
        return -1 if len(a) < len(b) else (1 if len(a) > len(b) else 0)
    return 0 if a is b else (-1 if id(a) < id(b) else 1)


def exact_eq(a, b):
    """=:= and pattern matching: 1 and 1.0 differ."""
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        return type(a) is type(b) and a == b
    if isinstance(a, Atom) or isinstance(b, Atom):
        return isinstance(a, Atom) and isinstance(b, Atom) and str(a) == str(b)
    if isinstance(a, tuple) and isinstance(b, tuple):
        return len(a) == len(b) and all(exact_eq(x, y) for x, y in zip(a, b))
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(exact_eq(x, y) for x, y in zip(a, b))
    if isinstance(a, bytes) and isinstance(b, bytes):
        return a == b
    return a is b


def erl_bool(v):
    if v == TRUE:
        return True
    if v == FALSE:
        return False
    raise ErlError("badarg: %r is not a boolean" % (v,))


def mk_bool(b):
    return TRUE if b else FALSE


# --------------------------------------------------------------------------- tokens
_TOKEN = re.compile(r"""
    (?P<ws>\s+|%[^\n]*)
  | (?P<char>\$(?:\\.|[^\\\s]))
  | (?P<bint>\d+\#[0-9a-zA-Z]+)
  | (?P<float>\d+\.\d+(?:[eE][+-]?\d+)?)
  | (?P<int>\d+)
  | (?P<var>[A-Z_][A-Za-z0-9_@]*)
  | (?P<atom>[a-z][A-Za-z0-9_@]*)
  | '(?P<qatom>(?:[^'\\]|\\.)*)'
  | "(?P<string>(?:[^"\\]|\\.)*)"
  | (?P<op><<|>>|<=|->|<-|\|\||=:=|=/=|==|/=|=<|>=|\+\+|--|[()\[\]{},;.|#=<>+\-*/!:?])
""", re.X)

KEYWORDS = {"case", "of", "end", "if", "fun", "when", "receive", "after", "begin", "and", "or", "not", "div", "rem",
            "andalso", "orelse", "band", "bor", "bxor", "bnot", "bsl", "bsr", "xor", "try", "catch"}


def tokenize(text):
    out, pos, line = [], 0, 1
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise ErlError("cannot tokenise at line %d: %r" % (line, text[pos:pos + 30]))
        kind = m.lastgroup
        s = m.group(0)
        if kind == "char":
            esc = {"s": " ", "n": "\n", "t": "\t", "\\": "\\", "r": "\r"}
            ch = esc.get(s[2], s[2]) if s[1] == "\\" else s[1]
            out.append(("int", ord(ch), line))
        elif kind == "bint":
            base, digits = s.split("#")
            out.append(("int", int(digits, int(base)), line))
        elif kind == "float":
            out.append(("float", float(s), line))
        elif kind == "int":
            out.append(("int", int(s), line))
        elif kind == "var":
            out.append(("var", s, line))
        elif kind == "atom":
            out.append(("kw", s, line) if s in KEYWORDS else ("atom", s, line))
        elif kind == "qatom":
            out.append(("atom", m.group("qatom"), line))
        elif kind == "string":
            raw = m.group("string")
            raw = raw.replace("\\n", "\n").replace("\\t", "\t").replace('\\"', '"').replace("\\\\", "\\")
            out.append(("string", raw, line))
        elif kind == "op":
            out.append(("op", s, line))
        line += s.count("\n")
        pos = m.end()
    out.append(("eof", None, line))
    return out


# --------------------------------------------------------------------------- parser -> AST (tuples)
class Parser:
    def __init__(self, tokens):
        self.t = tokens
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def at(self, kind, val=None):
        tok = self.t[self.i]
        return tok[0] == kind and (val is None or tok[1] == val)

    def accept(self, kind, val=None):
        if self.at(kind, val):
            return self.next()
        return None

    def expect(self, kind, val=None):
        tok = self.next()
        if tok[0] != kind or (val is not None and tok[1] != val):
            raise ErlError("line %d: expected %s %r, got %r" % (tok[2], kind, val, tok[:2]))
        return tok

    # ---- forms
    def forms(self):
        out = []
        while not self.at("eof"):
            out.append(self.form())
        return out

    def form(self):
        if self.at("op", "-"):
            self.next()
            name = self.expect("atom")[1]
            self.expect("op", "(")
            if name == "record":
                rec = self.expect("atom")[1]
                self.expect("op", ",")
                self.expect("op", "{")
                fields = []
                while not self.at("op", "}"):
                    f = self.expect("atom")[1]
                    default = None
                    if self.accept("op", "="):
                        default = self.expr()
                    fields.append((f, default))
                    self.accept("op", ",")
                self.expect("op", "}")
                self.expect("op", ")")
                self.expect("op", ".")
                return ("record", rec, fields)
            if name == "define":
                tok = self.next()
                self.expect("op", ",")
                body = self.expr()
                self.expect("op", ")")
                self.expect("op", ".")
                return ("define", tok[1], body)
            # -module, -export, ...: skip the balanced argument
            depth = 1
            while depth:
                tok = self.next()
                if tok[0] == "op" and tok[1] in "([{":
                    depth += 1
                elif tok[0] == "op" and tok[1] in ")]}":
                    depth -= 1
            self.expect("op", ".")
            return ("attribute", name)
        clauses = [self.fun_clause(True)]
        while self.accept("op", ";"):
            clauses.append(self.fun_clause(True))
        self.expect("op", ".")
        return ("function", clauses[0][0], len(clauses[0][1]), [c[1:] for c in clauses])

    def fun_clause(self, named):
        name = self.expect("atom")[1] if named else None
        line = self.peek()[2]
        self.expect("op", "(")
        args = []
        while not self.at("op", ")"):
            args.append(self.expr())
            self.accept("op", ",")
        self.expect("op", ")")
        guard = None
        if self.accept("kw", "when"):
            guard = self.guard()
        self.expect("op", "->")
        body = self.body()
        return (name, args, guard, body, line)

    def guard(self):
        # `,` = and, `;` = or  (a sequence of guard expressions)
        alts = []
        conj = [self.expr()]
        while True:
            if self.accept("op", ","):
                conj.append(self.expr())
            elif self.at("op", ";") and not self._clause_follows():
                self.next()
                alts.append(conj)
                conj = [self.expr()]
            else:
                break
        alts.append(conj)
        return alts

    def _clause_follows(self):
        return False

    def body(self):
        exprs = [self.expr()]
        while self.accept("op", ","):
            exprs.append(self.expr())
        return exprs

    # ---- expressions, by Erlang's operator priorities
    def expr(self):
        left = self.expr_orelse()
        if self.at("op", "=") or self.at("op", "!"):
            op = self.next()[1]
            right = self.expr()
            return ("match" if op == "=" else "send", left, right)
        return left

    def expr_orelse(self):
        left = self.expr_andalso()
        if self.accept("kw", "orelse"):
            return ("orelse", left, self.expr_orelse())
        return left

    def expr_andalso(self):
        left = self.expr_cmp()
        if self.accept("kw", "andalso"):
            return ("andalso", left, self.expr_andalso())
        return left

    def expr_cmp(self):
        left = self.expr_list()
        tok = self.peek()
        if tok[0] == "op" and tok[1] in ("==", "/=", "=<", "<", ">=", ">", "=:=", "=/="):
            self.next()
            return ("binop", tok[1], left, self.expr_list())
        return left

    def expr_list(self):
        left = self.expr_add()
        tok = self.peek()
        if tok[0] == "op" and tok[1] in ("++", "--"):
            self.next()
            return ("binop", tok[1], left, self.expr_list())
        return left

    def expr_add(self):
        left = self.expr_mul()
        while True:
            tok = self.peek()
            if (tok[0] == "op" and tok[1] in ("+", "-")) or (tok[0] == "kw" and tok[1] in ("or", "xor", "bor", "bxor", "bsl", "bsr")):
                self.next()
                left = ("binop", tok[1], left, self.expr_mul())
            else:
                return left

    def expr_mul(self):
        left = self.expr_unary()
        while True:
            tok = self.peek()
            if (tok[0] == "op" and tok[1] in ("*", "/")) or (tok[0] == "kw" and tok[1] in ("div", "rem", "and", "band")):
                self.next()
                left = ("binop", tok[1], left, self.expr_unary())
            else:
                return left

    def expr_unary(self):
        tok = self.peek()
        if (tok[0] == "op" and tok[1] in ("-", "+")) or (tok[0] == "kw" and tok[1] in ("not", "bnot")):
            self.next()
            return ("unop", tok[1], self.expr_unary())
        return self.expr_postfix()

    def expr_postfix(self):
        e = self.expr_primary()
        while True:
            if self.at("op", "#"):
                # Expr#rec.field  |  Expr#rec{updates}
                self.next()
                rec = self.expect("atom")[1]
                if self.accept("op", "."):
                    e = ("rec_get", e, rec, self.expect("atom")[1])
                else:
                    e = ("rec_update", e, rec, self.rec_fields())
            elif self.at("op", "("):
                self.next()
                args = []
                while not self.at("op", ")"):
                    args.append(self.expr())
                    self.accept("op", ",")
                self.expect("op", ")")
                e = ("call", e, args)
            elif self.at("op", ":") and e[0] == "atom" and self.peek(1)[0] == "atom":
                self.next()
                e = ("remote", e[1], self.expect("atom")[1])
            else:
                return e

    def rec_fields(self):
        self.expect("op", "{")
        fields = []
        while not self.at("op", "}"):
            f = self.next()[1]
            self.expect("op", "=")
            fields.append((f, self.expr()))
            self.accept("op", ",")
        self.expect("op", "}")
        return fields

    def expr_primary(self):
        tok = self.next()
        kind, val, line = tok
        if kind == "int" or kind == "float":
            return ("lit", val)
        if kind == "string":
            return ("lit", ErlString(ord(c) for c in val))
        if kind == "atom":
            return ("atom", val)
        if kind == "var":
            return ("var", val)
        if kind == "op":
            if val == "(":
                e = self.expr()
                self.expect("op", ")")
                return ("paren", e)
            if val == "{":
                items = []
                while not self.at("op", "}"):
                    items.append(self.expr())
                    self.accept("op", ",")
                self.expect("op", "}")
                return ("tuple", items)
            if val == "[":
                if self.accept("op", "]"):
                    return ("nil",)
                first = self.expr()
                if self.accept("op", "||"):
                    quals = []
                    while True:
                        save = self.i
                        pat = self.expr()
                        if self.accept("op", "<-"):
                            quals.append(("gen", pat, self.expr()))
                        elif self.accept("op", "<="):
                            quals.append(("bgen", pat, self.expr()))
                        else:
                            self.i = save
                            quals.append(("filter", self.expr()))
                        if not self.accept("op", ","):
                            break
                    self.expect("op", "]")
                    return ("lc", first, quals)
                items = [first]
                tail = ("nil",)
                while True:
                    if self.accept("op", ","):
                        items.append(self.expr())
                    elif self.accept("op", "|"):
                        tail = self.expr()
                        break
                    else:
                        break
                self.expect("op", "]")
                return ("cons", items, tail)
            if val == "#":
                rec = self.expect("atom")[1]
                if self.accept("op", "."):
                    return ("rec_index", rec, self.expect("atom")[1])
                return ("rec_new", rec, self.rec_fields())
            if val == "?":
                return ("macro", self.next()[1])
            if val == "<<":
                segs = []
                while not self.at("op", ">>"):
                    value = self.expr_primary()
                    size, typ = None, "integer"
                    if self.accept("op", ":"):
                        size = self.expect("int")[1]
                    if self.accept("op", "/"):
                        specs = [self.expect("atom")[1]]
                        while self.accept("op", "-"):
                            specs.append(self.expect("atom")[1])
                        typ = "-".join(specs)
                    segs.append((value, size, typ))
                    self.accept("op", ",")
                self.expect("op", ">>")
                return ("bin", segs)
        if kind == "kw":
            if val == "case":
                subject = self.expr()
                self.expect("kw", "of")
                clauses = self.cr_clauses()
                self.expect("kw", "end")
                return ("case", subject, clauses)
            if val == "if":
                clauses = []
                while True:
                    g = self.guard()
                    self.expect("op", "->")
                    clauses.append((g, self.body()))
                    if not self.accept("op", ";"):
                        break
                self.expect("kw", "end")
                return ("if", clauses)
            if val == "begin":
                b = self.body()
                self.expect("kw", "end")
                return ("block", b)
            if val == "receive":
                clauses = self.cr_clauses()
                self.expect("kw", "end")
                return ("receive", clauses)
            if val == "fun":
                if self.at("op", "("):
                    clauses = [self.fun_clause(False)]
                    while self.accept("op", ";"):
                        clauses.append(self.fun_clause(False))
                    self.expect("kw", "end")
                    return ("fun", [c[1:] for c in clauses])
                a = self.expect("atom")[1]
                if self.accept("op", ":"):
                    f = self.expect("atom")[1]
                    self.expect("op", "/")
                    return ("fun_ref", a, f, self.expect("int")[1])
                self.expect("op", "/")
                return ("fun_ref", None, a, self.expect("int")[1])
        raise ErlError("line %d: unexpected token %r" % (line, tok[:2]))

    def cr_clauses(self):
        clauses = []
        while True:
            pat = self.expr()
            guard = None
            if self.accept("kw", "when"):
                guard = self.guard()
            self.expect("op", "->")
            clauses.append((pat, guard, self.body()))
            if not self.accept("op", ";"):
                break
        return clauses


# --------------------------------------------------------------------------- evaluator
class Module:
    def __init__(self, text, name="raytracer"):
        self.name = name
        self.records = {}
        self.macros = {"MODULE": ("atom", name)}
        self.externals = {}    # (name, arity) -> Python callable standing in for a NIF (tests)
        self._pids = [(Atom("pid"), 0)]             # sequential process model: the running process is the last one
        self._next_pid = 1
        self.mailboxes = {self._pids[0]: []}
        self.functions = {}
        self.out = []          # io:format output (stdout)
        self.files = {}        # filename -> list of written chunks
        self.calls = 0
        for f in Parser(tokenize(text)).forms():
            if f[0] == "record":
                self.records[f[1]] = f[2]
            elif f[0] == "define":
                self.macros[f[1]] = f[2]
            elif f[0] == "function":
                self.functions[(f[1], f[2])] = f[3]

    # ---- public
    def call(self, fname, *args):
        return self.apply_local(fname, list(args))

    def apply_local(self, fname, args):
        ext = self.externals.get((fname, len(args)))
        if ext is not None:
            return ext(*args)
        clauses = self.functions.get((fname, len(args)))
        if clauses is None:
            raise ErlError("undef: %s/%d" % (fname, len(args)))
        self.calls += 1
        for pats, guard, body, _line in clauses:
            env = {}
            if all(self.match(p, a, env) for p, a in zip(pats, args)) and self.guard_ok(guard, env):
                return self.eval_body(body, env)
        raise ErlError("function_clause: %s/%d %r" % (fname, len(args), args))

    # ---- records
    def rec_index(self, rec, field):
        for k, (f, _d) in enumerate(self.records[rec]):
            if f == field:
                return k + 1
        raise ErlError("record %s has no field %s" % (rec, field))

    # ---- patterns
    def match(self, pat, val, env):
        k = pat[0]
        if k == "var":
            name = pat[1]
            if name == "_":
                return True
            if name in env:
                return exact_eq(env[name], val)
            env[name] = val
            return True
        if k == "lit":
            return exact_eq(pat[1], val)
        if k == "atom":
            return isinstance(val, Atom) and str(val) == pat[1]
        if k == "tuple":
            return (isinstance(val, tuple) and len(val) == len(pat[1])
                    and all(self.match(p, v, env) for p, v in zip(pat[1], val)))
        if k == "nil":
            return isinstance(val, list) and len(val) == 0
        if k == "cons":
            items, tail = pat[1], pat[2]
            if not isinstance(val, list) or len(val) < len(items):
                return False
            for p, v in zip(items, val):
                if not self.match(p, v, env):
                    return False
            return self.match(tail, val[len(items):], env)
        if k == "rec_new":
            rec, fields = pat[1], pat[2]
            decl = self.records[rec]
            if not (isinstance(val, tuple) and len(val) == len(decl) + 1 and isinstance(val[0], Atom) and str(val[0]) == rec):
                return False
            for f, p in fields:
                if not self.match(p, val[self.rec_index(rec, f)], env):
                    return False
            return True
        if k == "match":
            return self.match(pat[1], val, env) and self.match(pat[2], val, env)
        if k == "paren":
            return self.match(pat[1], val, env)
        if k == "unop" and pat[1] == "-" and pat[2][0] == "lit":
            return exact_eq(-pat[2][1], val)
        if k == "macro":
            return self.match(self.macros[pat[1]], val, env)
        if k == "bin":
            # <<Var:Size/Type, ...>> against a binary: whole-byte float and unsigned big-endian integer segments
            if not isinstance(val, bytes):
                return False
            pos = 0
            for value, size, typ in pat[1]:
                kind, order = self.seg_type(typ)
                if kind == "binary" and size is None:
                    chunk = val[pos:]                # Rest/binary: the tail
                    pos = len(val)
                    if not self.match(value, chunk, env):
                        return False
                    continue
                size = (64 if kind == "float" else 8) if size is None else size
                if kind == "binary":
                    size *= 8
                if size % 8 or pos + size // 8 > len(val):
                    return False
                chunk = val[pos:pos + size // 8]
                pos += size // 8
                if kind == "float":
                    if size not in (32, 64):
                        return False
                    v = struct.unpack(order + ("f" if size == 32 else "d"), chunk)[0]
                    if math.isinf(v) or math.isnan(v):
                        return False                 # BEAM does not match non-finite floats
                elif kind == "integer":
                    v = int.from_bytes(chunk, "big" if order == ">" else "little")
                else:
                    v = chunk
                if not self.match(value, v, env):
                    return False
            return pos == len(val)
        raise ErlError("unsupported pattern %r" % (pat,))

    def guard_ok(self, guard, env):
        if guard is None:
            return True
        for conj in guard:
            ok = True
            for g in conj:
                try:
                    v = self.eval(g, env)
                except ErlError:
                    v = FALSE              # an exception in a guard makes it fail
                if v != TRUE:
                    ok = False
                    break
            if ok:
                return True
        return False

    # ---- expressions
    def eval_body(self, body, env):
        v = None
        for e in body:
            v = self.eval(e, env)
        return v

    def eval(self, e, env):
        k = e[0]
        if k == "lit":
            return e[1]
        if k == "var":
            try:
                return env[e[1]]
            except KeyError:
                raise ErlError("unbound variable %s" % e[1])
        if k == "atom":
            return Atom(e[1])
        if k == "call":
            return self.eval_call(e, env)
        if k == "binop":
            return self.binop(e[1], e[2], e[3], env)
        if k == "rec_get":
            v = self.eval(e[1], env)
            rec = e[2]
            if not (isinstance(v, tuple) and v and isinstance(v[0], Atom) and str(v[0]) == rec and len(v) == len(self.records[rec]) + 1):
                raise ErlError("badrecord %s: %r" % (rec, v))
            return v[self.rec_index(rec, e[3])]
        if k == "rec_new":
            rec = e[1]
            given = {f: x for f, x in e[2]}
            vals = [Atom(rec)]
            for f, default in self.records[rec]:
                if f in given:
                    vals.append(self.eval(given[f], env))
                elif default is not None:
                    vals.append(self.eval(default, env))
                else:
                    vals.append(Atom("undefined"))
            return tuple(vals)
        if k == "rec_update":
            v = list(self.eval(e[1], env))
            for f, x in e[3]:
                v[self.rec_index(e[2], f)] = self.eval(x, env)
            return tuple(v)
        if k == "rec_index":
            return self.rec_index(e[1], e[2]) + 1
        if k == "paren":
            return self.eval(e[1], env)
        if k == "unop":
            v = self.eval(e[2], env)
            if e[1] == "-":
                self.need_num(v)
                return -v
            if e[1] == "+":
                self.need_num(v)
                return v
            if e[1] == "not":
                return mk_bool(not erl_bool(v))
            raise ErlError("unsupported unary %s" % e[1])
        if k == "tuple":
            return tuple(self.eval(x, env) for x in e[1])
        if k == "nil":
            return []
        if k == "cons":
            head = [self.eval(x, env) for x in e[1]]
            tail = self.eval(e[2], env)
            if not isinstance(tail, list):
                raise ErlError("improper lists are not supported")
            return head + tail
        if k == "match":
            v = self.eval(e[2], env)
            if not self.match(e[1], v, env):
                raise ErlError("badmatch: %r" % (v,))
            return v
        if k == "case":
            v = self.eval(e[1], env)
            for pat, guard, body in e[2]:
                trial = dict(env)
                if self.match(pat, v, trial) and self.guard_ok(guard, trial):
                    env.update(trial)
                    return self.eval_body(body, env)
            raise ErlError("case_clause: %r" % (v,))
        if k == "if":
            for guard, body in e[1]:
                if self.guard_ok(guard, env):
                    return self.eval_body(body, env)
            raise ErlError("if_clause")
        if k == "block":
            return self.eval_body(e[1], env)
        if k == "fun":
            clauses = e[1]
            closure = dict(env)
            arity = len(clauses[0][0])

            def fun(*args):
                if len(args) != arity:
                    raise ErlError("badarity")
                for pats, guard, body, _line in clauses:
                    local = dict(closure)
                    # fun-clause heads shadow nothing that is bound in the closure only when equal (Erlang
                    # would warn about shadowing; raytracer.erl has no such case)
                    if all(self.match(p, a, local) for p, a in zip(pats, args)) and self.guard_ok(guard, local):
                        return self.eval_body(body, local)
                raise ErlError("function_clause in fun: %r" % (args,))
            return fun
        if k == "fun_ref":
            mod, fname, arity = e[1], e[2], e[3]
            if mod is not None and mod != self.name:
                return lambda *a: self.apply_remote(mod, fname, list(a))
            return lambda *a: self.apply_local(fname, list(a))
        if k == "lc":
            out = []
            self.lc(e[1], e[2], 0, dict(env), out)
            return out
        if k == "macro":
            return self.eval(self.macros[e[1]], env)
        if k == "bin":
            out = b""
            for value, size, typ in e[1]:
                v = self.eval(value, env)
                kind, order = self.seg_type(typ)
                if kind == "binary":
                    if not isinstance(v, bytes):
                        raise ErlError("badarg: binary segment")
                    out += v if size is None else v[:size]
                    continue
                size = (64 if kind == "float" else 8) if size is None else size
                if kind == "float":
                    self.need_num(v)
                    if size not in (32, 64):
                        raise ErlError("badarg: float segment of %d bits" % size)
                    try:
                        out += struct.pack(order + ("f" if size == 32 else "d"), float(v))   # round to nearest even, like BEAM
                    except OverflowError:
                        raise ErlError("badarg: float does not fit the segment")
                else:
                    if not isinstance(v, int) or size % 8:
                        raise ErlError("badarg: integer segment")
                    out += (v & ((1 << size) - 1)).to_bytes(size // 8, "big" if order == ">" else "little")
            return out
        if k == "andalso":
            return self.eval(e[2], env) if erl_bool(self.eval(e[1], env)) else FALSE
        if k == "orelse":
            return TRUE if erl_bool(self.eval(e[1], env)) else self.eval(e[2], env)
        if k == "send":
            # Pid ! Msg in the sequential process model below
            pid = self.eval(e[1], env)
            msg = self.eval(e[2], env)
            if pid not in self.mailboxes:
                raise ErlError("badarg: send to %r" % (pid,))
            self.mailboxes[pid].append(msg)
            return msg
        if k == "receive":
            # selective receive from the current process's mailbox: first message (in arrival order) that a
            # clause accepts.  A process that would have to WAIT is an error here: spawned funs run to completion
            # at the spawn (one valid schedule of programs whose children never wait for their parent).
            box = self.mailboxes[self._pids[-1]]
            for n, msg in enumerate(box):
                for pat, guard, body in e[1]:
                    local = dict(env)
                    if self.match(pat, msg, local) and self.guard_ok(guard, local):
                        del box[n]
                        env.update(local)
                        return self.eval_body(body, env)
            raise ErlError("receive would block: no matching message in %r" % (box,))
        raise ErlError("unsupported expression %r" % (k,))

    def lc(self, template, quals, qi, env, out):
        if qi == len(quals):
            out.append(self.eval(template, env))
            return
        q = quals[qi]
        if q[0] == "gen":
            for item in self.eval(q[2], env):
                local = dict(env)
                if self.match(q[1], item, local):
                    self.lc(template, quals, qi + 1, local, out)
        elif q[0] == "bgen":
            data = self.eval(q[2], env)
            if not isinstance(data, bytes) or q[1][0] != "bin":
                raise ErlError("bad binary generator")
            width = sum(((64 if self.seg_type(t)[0] == "float" else 8) if sz is None else sz) for _v, sz, t in q[1][1]) // 8
            for k in range(0, len(data) - width + 1, width):
                local = dict(env)
                if self.match(q[1], data[k:k + width], local):
                    self.lc(template, quals, qi + 1, local, out)
        else:
            if erl_bool(self.eval(q[1], env)):
                self.lc(template, quals, qi + 1, env, out)

    @staticmethod
    def seg_type(typ):
        """'float-native' -> ('float', byte order char for struct); default big-endian like Erlang."""
        parts = typ.split("-")
        kind = next((p for p in parts if p in ("integer", "float", "binary")), "integer")
        order = "<" if ("little" in parts or ("native" in parts and sys.byteorder == "little")) else ">"
        return kind, order

    @staticmethod
    def need_num(v):
        if isinstance(v, bool) or not isinstance(v, (int, float)):
            raise ErlError("badarith: %r" % (v,))

    def binop(self, op, le, re_, env):
        a = self.eval(le, env)
        b = self.eval(re_, env)            # `and` / `or` are strict: both sides are evaluated (erl:319)
        if op in ("+", "-", "*"):
            self.need_num(a); self.need_num(b)
            if isinstance(a, float) or isinstance(b, float):
                a, b = float(a), float(b)
            r = a + b if op == "+" else (a - b if op == "-" else a * b)
            if isinstance(r, float) and (math.isinf(r) or math.isnan(r)):
                raise ErlError("badarith: float overflow")
            return r
        if op == "/":
            self.need_num(a); self.need_num(b)
            if float(b) == 0.0:
                raise ErlError("badarith: division by zero")
            r = float(a) / float(b)
            if math.isinf(r) or math.isnan(r):
                raise ErlError("badarith: float overflow")
            return r
        if op == "div" or op == "rem":
            if not (isinstance(a, int) and isinstance(b, int)) or b == 0:
                raise ErlError("badarith")
            q = abs(a) // abs(b) * (1 if (a >= 0) == (b >= 0) else -1)
            return q if op == "div" else a - b * q
        if op == "==":
            return mk_bool(term_cmp(a, b) == 0)
        if op == "/=":
            return mk_bool(term_cmp(a, b) != 0)
        if op == "<":
            return mk_bool(term_cmp(a, b) < 0)
        if op == ">":
            return mk_bool(term_cmp(a, b) > 0)
        if op == "=<":
            return mk_bool(term_cmp(a, b) <= 0)
        if op == ">=":
            return mk_bool(term_cmp(a, b) >= 0)
        if op == "=:=":
            return mk_bool(exact_eq(a, b))
        if op == "=/=":
            return mk_bool(not exact_eq(a, b))
        if op in ("band", "bor", "bxor", "bsl", "bsr"):
            if not (isinstance(a, int) and isinstance(b, int)) or isinstance(a, bool) or isinstance(b, bool):
                raise ErlError("badarith: %r %s %r" % (a, op, b))
            return {"band": a & b, "bor": a | b, "bxor": a ^ b, "bsl": a << b if op == "bsl" else 0,
                    "bsr": a >> b if op == "bsr" else 0}[op]
        if op == "and":
            return mk_bool(erl_bool(a) and erl_bool(b))
        if op == "or":
            return mk_bool(erl_bool(a) or erl_bool(b))
        if op == "xor":
            return mk_bool(erl_bool(a) != erl_bool(b))
        if op == "++":
            return list(a) + list(b)
        raise ErlError("unsupported operator %s" % op)

    # ---- calls
    def eval_call(self, e, env):
        target, arg_exprs = e[1], e[2]
        args = [self.eval(x, env) for x in arg_exprs]
        if target[0] == "atom":
            fname = target[1]
            if (fname, len(args)) in self.functions:
                return self.apply_local(fname, args)
            return self.bif(fname, args)
        if target[0] == "remote":
            return self.apply_remote(target[1], target[2], args)
        f = self.eval(target, env)
        if not callable(f):
            raise ErlError("badfun: %r" % (f,))
        return f(*args)

    def bif(self, fname, args):
        if fname == "trunc":
            self.need_num(args[0])
            return int(args[0])
        if fname == "abs":
            self.need_num(args[0])
            return abs(args[0])
        if fname == "length":
            return len(args[0])
        if fname == "float":
            return float(args[0])
        if fname == "round":
            v = args[0]
            return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))
        if fname == "hd":
            return args[0][0]
        if fname == "tl":
            return args[0][1:]
        if fname == "element":
            return args[1][args[0] - 1]
        if fname == "list_to_integer":
            return int(args[0].text() if isinstance(args[0], ErlString) else "".join(chr(c) for c in args[0]))
        if fname == "list_to_atom":
            return Atom(args[0].text() if isinstance(args[0], ErlString) else "".join(chr(c) for c in args[0]))
        if fname in ("is_list", "is_atom", "is_tuple", "is_binary", "is_number"):
            kinds = {"is_list": list, "is_atom": Atom, "is_tuple": tuple, "is_binary": bytes, "is_number": (int, float)}
            return mk_bool(isinstance(args[0], kinds[fname]) and not isinstance(args[0], bool))
        if fname == "exit":
            raise ErlExit(args[0])
        if fname == "self" and not args:
            return self._pids[-1]
        if fname in ("spawn", "spawn_link") and len(args) == 1:
            pid = (Atom("pid"), self._next_pid)
            self._next_pid += 1
            self.mailboxes[pid] = []
            self._pids.append(pid)
            try:
                args[0]()                            # runs to completion here; an exit of a linked child reaches the parent
            except ErlExit:
                if fname == "spawn_link":
                    raise
            finally:
                self._pids.pop()
            return pid
        if fname == "integer_to_list":
            return ErlString(ord(c) for c in str(args[0]))
        if fname == "min":
            return args[0] if term_cmp(args[0], args[1]) <= 0 else args[1]
        if fname == "max":
            return args[0] if term_cmp(args[0], args[1]) >= 0 else args[1]
        if fname == "is_integer":
            return mk_bool(isinstance(args[0], int))
        if fname == "is_float":
            return mk_bool(isinstance(args[0], float))
        raise ErlError("undef: %s/%d" % (fname, len(args)))

    def apply_remote(self, mod, fname, args):
        if mod == self.name:
            return self.apply_local(fname, args)
        if mod == "math":
            try:
                if fname == "pi":
                    return math.pi
                if fname == "sqrt":
                    return math.sqrt(args[0])
                if fname == "pow":
                    return math.pow(args[0], args[1])
                if fname == "tan":
                    return math.tan(args[0])
                if fname == "sin":
                    return math.sin(args[0])
                if fname == "cos":
                    return math.cos(args[0])
            except (ValueError, OverflowError):
                raise ErlError("badarith in math:%s%r" % (fname, tuple(args)))
        if mod == "lists":
            if fname == "foldl":
                f, acc, lst = args
                for x in lst:
                    acc = f(x, acc)
                return acc
            if fname == "map":
                return [args[0](x) for x in args[1]]
            if fname == "flatmap":
                out = []
                for x in args[1]:
                    out.extend(args[0](x))
                return out
            if fname == "foreach":
                for x in args[1]:
                    args[0](x)
                return Atom("ok")
            if fname == "seq":
                return list(range(args[0], args[1] + 1))
            if fname == "max":
                best = args[0][0]
                for x in args[0][1:]:
                    if term_cmp(x, best) > 0:          # lists:max keeps the first of equal elements
                        best = x
                return best
            if fname == "min":
                best = args[0][0]
                for x in args[0][1:]:
                    if term_cmp(x, best) < 0:
                        best = x
                return best
            if fname == "keysort":
                import functools
                n = args[0]
                return sorted(args[1], key=functools.cmp_to_key(lambda p, q: term_cmp(p[n - 1], q[n - 1])))
            if fname == "split":
                return (args[1][:args[0]], args[1][args[0]:])
            if fname == "reverse":
                return list(reversed(args[0]))
            if fname == "zip":
                if len(args[0]) != len(args[1]):
                    raise ErlError("function_clause: lists:zip/2 of unequal lists")
                return [(a, b) for a, b in zip(args[0], args[1])]
        if mod == "io" and fname == "format":
            if len(args) == 1:
                self.out.append(self.format(args[0], []))
            elif len(args) == 2:
                self.out.append(self.format(args[0], args[1]))
            else:
                dev = args[0]
                self.files[dev[1]].append(self.format(args[1], args[2]))
            return Atom("ok")
        if mod == "erlang" and fname == "nif_error":
            raise ErlError("nif_error: %r" % (args[0],))
        if mod == "erlang" and fname in ("min", "max", "exit", "is_list", "integer_to_list"):
            return self.bif(fname, args)
        if mod == "proplists" and fname == "get_value":
            for item in args[1]:
                if isinstance(item, tuple) and len(item) == 2 and exact_eq(item[0], args[0]):
                    return item[1]
            return args[2] if len(args) > 2 else Atom("undefined")
        if mod == "io_lib" and fname == "format":
            return ErlString(ord(c) for c in self.format(args[0], args[1]))
        if mod == "file" and fname == "write_file":
            def flat(x):
                if isinstance(x, bytes):
                    return x.decode("latin-1")
                if isinstance(x, int):
                    return chr(x)
                return "".join(flat(y) for y in x)
            name = args[0].text() if isinstance(args[0], ErlString) else "".join(chr(c) for c in args[0])
            self.files[name] = [flat(args[1])]
            return Atom("ok")
        if mod == "file":
            if fname == "open":
                name = args[0].text() if isinstance(args[0], ErlString) else str(args[0])
                self.files[name] = []
                return (Atom("ok"), (Atom("io_device"), name))
            if fname == "close":
                return Atom("ok")
        raise ErlError("undef: %s:%s/%d" % (mod, fname, len(args)))

    # io:format with the directives raytracer.erl uses: ~p ~w ~n ~s
    def format(self, fmt, args):
        text = fmt.text() if isinstance(fmt, ErlString) else "".join(chr(c) for c in fmt)
        out, k, i = [], 0, 0
        while i < len(text):
            c = text[i]
            if c == "~" and i + 1 < len(text):
                d = text[i + 1]
                i += 2
                if d == "n":
                    out.append("\n")
                elif d in "pw":
                    out.append(self.show(args[k])); k += 1
                elif d == "s":
                    a = args[k]; k += 1
                    out.append(a.text() if isinstance(a, ErlString) else "".join(chr(x) for x in a))
                else:
                    raise ErlError("io:format directive ~%s is not supported" % d)
            else:
                out.append(c)
                i += 1
        return "".join(out)

    def show(self, v):
        if isinstance(v, Atom):
            return str.__str__(v)
        if isinstance(v, int):
            return str(v)
        if isinstance(v, float):
            return repr(v)                 # shortest round-trip digits, like OTP's ~p (exponent form differs)
        if isinstance(v, tuple):
            return "{" + ",".join(self.show(x) for x in v) + "}"
        if isinstance(v, ErlString):
            return '"' + v.text() + '"'
        if isinstance(v, list):
            return "[" + ",".join(self.show(x) for x in v) + "]"
        return "#Fun"


def load(path="/root/reference/raytracer.erl"):
    with open(path) as fh:
        return Module(fh.read())


# --------------------------------------------------------------------------- term builders for tests
def vec(x, y, z):
    return (Atom("vector"), x, y, z)


def colour(r, g, b):
    return (Atom("colour"), r, g, b)


def material(col, sp, sh, refl):
    return (Atom("material"), colour(*col), sp, sh, refl)


def to_py(v):
    """Erlang term -> plain Python (atoms become str) for JSON."""
    if isinstance(v, Atom):
        return str.__str__(v)
    if isinstance(v, (tuple, list)):
        return [to_py(x) for x in v]
    return v


if __name__ == "__main__":
    sys.setrecursionlimit(100000)
    m = load(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/raytracer.erl")
    print("functions:", len(m.functions), "records:", sorted(m.records))
    print("run_tests() ->", m.call("run_tests"))
    print("".join(m.out))

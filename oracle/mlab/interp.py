"""A small interpreter for the MATLAB subset the reference's RK4 path is written in.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Neither MATLAB nor Octave exists in this image,
so the reference's own .m files cannot be run by their native interpreter.  This module
executes them UNMODIFIED, straight from /root/reference (classdef handle classes,
methods, scripts, nested indexing with `end`, matrix literals, nargin/nargout, ...), so
that golden vectors under tests/golden/ come from the reference's source text -- its loop
structure, index arithmetic and operation order -- rather than from a hand restatement.

What is NOT the reference here: MATLAB's built-in functions.  Those are restated in
oracle/mlab/builtins.py (linspace rounding, mtimes summation order, exp,
griddedInterpolant's linear formula) and documented there.

Supported: scripts, function files, classdef files (properties / methods blocks, abstract
blocks are skipped), if/elseif/else, for, while, break/continue/return, multiple return
values with ~ placeholders, anonymous functions, struct fields, handle-object properties
and methods (both obj.method(...) and method(obj, ...)), N-d indexing and indexed
assignment with `end` and `:`, command-syntax no-ops (close all, clear all, clc).
"""
from __future__ import annotations

import os
import re

import numpy as np

from . import builtins as B

# ------------------------------------------------------------------------------ tokens
_TOKEN_RE = re.compile(r"""
    (?P<num>(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?))
  | (?P<id>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op>\.\^|\.\*|\./|\.'|==|~=|<=|>=|&&|\|\||[-+*/\\^<>=&|~:,;()\[\]{}@.'])
""", re.X)

KEYWORDS = {"if", "elseif", "else", "end", "for", "while", "function", "return", "break", "continue",
            "classdef", "properties", "methods", "switch", "case", "otherwise"}


class Tok:
    __slots__ = ("kind", "val", "sp", "line")

    def __init__(self, kind, val, sp, line):
        self.kind, self.val, self.sp, self.line = kind, val, sp, line

    def __repr__(self):
        return f"{self.kind}:{self.val!r}"


def tokenize(src: str):
    toks = []
    i, n, line = 0, len(src), 1
    sp = False            # whitespace before the next token
    depth = 0             # bracket depth ([, {) for quote disambiguation
    while i < n:
        c = src[i]
        if c in " \t\r":
            sp = True
            i += 1
            continue
        if c == "\n":
            toks.append(Tok("nl", "\n", sp, line))
            line += 1
            i += 1
            sp = False
            continue
        if c == "%":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if src.startswith("...", i):
            while i < n and src[i] != "\n":
                i += 1
            i += 1
            line += 1
            sp = True
            continue
        if c == "'":
            prev = toks[-1] if toks else None
            is_transpose = (prev is not None and not sp and
                            (prev.kind in ("num", "id") and prev.val not in KEYWORDS or prev.val in (")", "]", "}", "'", ".'")))
            # inside brackets `[a 'str']` a quote after whitespace starts a string
            if prev is not None and sp and depth > 0:
                is_transpose = False
            elif prev is not None and sp and (prev.kind in ("num", "id") and prev.val not in KEYWORDS or prev.val in (")", "]", "}", "'")):
                is_transpose = True
            if is_transpose:
                toks.append(Tok("op", "'", sp, line))
                i += 1
                sp = False
                continue
            j = i + 1
            buf = []
            while j < n:
                if src[j] == "'":
                    if j + 1 < n and src[j + 1] == "'":
                        buf.append("'")
                        j += 2
                        continue
                    break
                buf.append(src[j])
                j += 1
            toks.append(Tok("str", "".join(buf), sp, line))
            i = j + 1
            sp = False
            continue
        m = _TOKEN_RE.match(src, i)
        if not m:
            raise SyntaxError(f"line {line}: cannot tokenize {src[i:i+20]!r}")
        if m.lastgroup == "num":
            toks.append(Tok("num", m.group("num"), sp, line))
        elif m.lastgroup == "id":
            toks.append(Tok("id", m.group("id"), sp, line))
        else:
            op = m.group("op")
            if op in "[{":
                depth += 1
            elif op in "]}":
                depth -= 1
            toks.append(Tok("op", op, sp, line))
        i = m.end()
        sp = False
    toks.append(Tok("nl", "\n", False, line))
    toks.append(Tok("eof", None, False, line))
    return toks


# ------------------------------------------------------------------------------ parser
class Parser:
    def __init__(self, toks, fname="<src>"):
        self.t = toks
        self.i = 0
        self.fname = fname
        self.idx_depth = 0      # > 0 while parsing inside index parentheses: `end` is a value there
        self.in_matrix = 0

    # -- helpers
    def peek(self, k=0):
        return self.t[self.i + k]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def at(self, val):
        tok = self.peek()
        return tok.kind in ("op", "id") and tok.val == val

    def expect(self, val):
        tok = self.next()
        if tok.val != val:
            raise SyntaxError(f"{self.fname}:{tok.line}: expected {val!r}, got {tok.val!r}")
        return tok

    def skip_nl(self):
        while self.peek().kind == "nl" or self.at(";") or self.at(","):
            self.next()

    # -- file level
    def parse_file(self):
        self.skip_nl()
        if self.at("classdef"):
            return ("classdef", self.parse_classdef())
        if self.at("function"):
            funcs = []
            while self.at("function"):
                funcs.append(self.parse_function())
                self.skip_nl()
            return ("functions", funcs)
        return ("script", self.parse_block(("eof",)))

    def parse_classdef(self):
        self.expect("classdef")
        name = self.next().val
        supers = []
        if self.at("<"):
            self.next()
            supers.append(self.next().val)
            while self.at("&"):
                self.next()
                supers.append(self.next().val)
        props, methods = [], {}
        self.skip_nl()
        while not self.at("end"):
            kind = self.next().val
            attrs = ""
            if self.at("("):
                depth = 0
                while True:
                    tok = self.next()
                    attrs += str(tok.val)
                    if tok.val == "(":
                        depth += 1
                    if tok.val == ")":
                        depth -= 1
                        if depth == 0:
                            break
            abstract = "Abstract" in attrs
            self.skip_nl()
            if kind == "properties":
                while not self.at("end"):
                    pname = self.next().val
                    default = None
                    if self.at("="):
                        self.next()
                        default = self.parse_expr()
                    props.append((pname, default))
                    self.skip_nl()
                self.expect("end")
            elif kind == "methods":
                if abstract:
                    # signatures only: skip to the block's closing `end` (signatures hold no blocks)
                    while not (self.at("end") and self.peek(-1).kind == "nl"):
                        self.next()
                    self.expect("end")
                else:
                    while not self.at("end"):
                        f = self.parse_function()
                        methods[f["name"]] = f
                        self.skip_nl()
                    self.expect("end")
            else:
                raise SyntaxError(f"{self.fname}: unexpected {kind!r} in classdef")
            self.skip_nl()
        self.expect("end")
        return dict(name=name, supers=supers, props=props, methods=methods)

    def parse_function(self):
        self.expect("function")
        outs = []
        # forms: function name(args) | function out = name(args) | function [o1,o2] = name(args)
        if self.at("["):
            self.next()
            while not self.at("]"):
                if self.at(","):
                    self.next()
                    continue
                outs.append(self.next().val)
            self.next()
            self.expect("=")
            name = self.next().val
        else:
            first = self.next().val
            if self.at("="):
                self.next()
                outs = [first]
                name = self.next().val
            else:
                name = first
        while self.at("."):           # e.g. set.Prop -- not used by the reference, tolerated
            self.next()
            name += "." + self.next().val
        params = []
        if self.at("("):
            self.next()
            while not self.at(")"):
                if self.at(","):
                    self.next()
                    continue
                params.append(self.next().val)   # may be '~'
            self.next()
        body, nested = [], {}
        self.skip_nl()
        while True:
            if self.at("end"):
                self.next()
                break
            if self.peek().kind == "eof":
                break
            if self.at("function"):
                # nested function (end-terminated files) -- or the next top-level function in a file
                # whose functions are not end-terminated; the reference only has the former
                f = self.parse_function()
                nested[f["name"]] = f
                self.skip_nl()
                continue
            body.append(self.parse_statement())
            self.skip_nl()
        return dict(name=name, outs=outs, params=params, body=body, nested=nested)

    # -- statements
    def parse_block(self, terminators):
        stmts = []
        self.skip_nl()
        while True:
            tok = self.peek()
            if tok.kind == "eof" and "eof" in terminators:
                break
            if tok.kind == "id" and tok.val in terminators:
                break
            if tok.kind == "eof":
                raise SyntaxError(f"{self.fname}: unexpected end of file")
            stmts.append(self.parse_statement())
            self.skip_nl()
        return stmts

    def parse_statement(self):
        tok = self.peek()
        if tok.kind == "id":
            v = tok.val
            if v == "if":
                return self.parse_if()
            if v == "for":
                self.next()
                paren = self.at("(")
                if paren:
                    self.next()
                var = self.next().val
                self.expect("=")
                rng = self.parse_expr()
                if paren:
                    self.expect(")")
                body = self.parse_block(("end",))
                self.expect("end")
                return ("for", var, rng, body)
            if v == "while":
                self.next()
                cond = self.parse_expr()
                body = self.parse_block(("end",))
                self.expect("end")
                return ("while", cond, body)
            if v in ("return", "break", "continue"):
                self.next()
                return (v,)
            # command syntax no-ops:  close all / clear all / clc / hold on ...
            if v in ("close", "clear", "clc", "hold", "format", "figure") and self.peek(1).kind in ("id", "nl"):
                nxt = self.peek(1)
                if nxt.kind == "nl" or (nxt.kind == "id" and self.peek(2).kind in ("nl", "eof") or self.peek(2).val == ";"):
                    while self.peek().kind != "nl":
                        self.next()
                    return ("nop",)
        # multi-assignment  [a, b] = f(...)
        if self.at("["):
            save = self.i
            try:
                lhs = self.parse_multi_lhs()
                if self.at("=") and not self.at("=="):
                    self.next()
                    rhs = self.parse_expr()
                    return ("massign", lhs, rhs)
            except SyntaxError:
                pass
            self.i = save
        expr = self.parse_expr()
        if self.at("="):
            self.next()
            rhs = self.parse_expr()
            return ("assign", self.to_lvalue(expr), rhs)
        return ("expr", expr)

    def parse_multi_lhs(self):
        self.expect("[")
        out = []
        while not self.at("]"):
            if self.at(","):
                self.next()
                continue
            if self.at("~"):
                self.next()
                out.append(None)
                continue
            self.idx_depth_save = self.idx_depth
            e = self.parse_postfix()
            out.append(self.to_lvalue(e))
        self.next()
        return out

    def to_lvalue(self, e):
        if e[0] in ("name", "index", "field", "dynfield", "cellindex"):
            return e
        raise SyntaxError(f"{self.fname}: invalid assignment target {e[0]}")

    def parse_if(self):
        self.expect("if")
        clauses = []
        cond = self.parse_expr()
        body = self.parse_block(("elseif", "else", "end"))
        clauses.append((cond, body))
        els = None
        while True:
            if self.at("elseif"):
                self.next()
                cond = self.parse_expr()
                body = self.parse_block(("elseif", "else", "end"))
                clauses.append((cond, body))
            elif self.at("else"):
                self.next()
                els = self.parse_block(("end",))
            else:
                break
        self.expect("end")
        return ("if", clauses, els)

    # -- expressions (precedence climbing, MATLAB's table)
    def parse_expr(self):
        return self.parse_oror()

    def _binleft(self, sub, ops):
        e = sub()
        while self.peek().kind == "op" and self.peek().val in ops and not self._matrix_break():
            op = self.next().val
            e = ("bin", op, e, sub())
        return e

    def _matrix_break(self):
        """Inside [ ], `a -b` (space before, none after) starts a new element."""
        if not self.in_matrix or self.idx_depth_in_matrix():
            return False
        tok, nxt = self.peek(), self.peek(1)
        return tok.val in ("+", "-") and tok.sp and not nxt.sp

    def idx_depth_in_matrix(self):
        return self._paren_in_matrix > 0

    _paren_in_matrix = 0

    def parse_oror(self):
        return self._binleft(self.parse_andand, ("||",))

    def parse_andand(self):
        return self._binleft(self.parse_or, ("&&",))

    def parse_or(self):
        return self._binleft(self.parse_and, ("|",))

    def parse_and(self):
        return self._binleft(self.parse_cmp, ("&",))

    def parse_cmp(self):
        return self._binleft(self.parse_range, ("<", "<=", ">", ">=", "==", "~="))

    def parse_range(self):
        # a:b  and  a:s:b ; a bare ':' inside an index is handled in parse_args
        first = self.parse_add()
        if self.at(":") and not self._colon_is_bare():
            self.next()
            second = self.parse_add()
            if self.at(":") and not self._colon_is_bare():
                self.next()
                third = self.parse_add()
                return ("range", first, second, third)
            return ("range", first, None, second)
        return first

    def _colon_is_bare(self):
        nxt = self.peek(1)
        return nxt.val in (",", ")") and nxt.kind == "op"

    def parse_add(self):
        return self._binleft(self.parse_mul, ("+", "-"))

    def parse_mul(self):
        return self._binleft(self.parse_unary, ("*", "/", ".*", "./", "\\"))

    def parse_unary(self):
        if self.peek().kind == "op" and self.peek().val in ("-", "+", "~"):
            op = self.next().val
            return ("un", op, self.parse_unary())
        return self.parse_power()

    def parse_power(self):
        base = self.parse_postfix()
        while self.peek().kind == "op" and self.peek().val in ("^", ".^"):
            op = self.next().val
            # exponent may carry a unary sign: 2^-1
            if self.peek().kind == "op" and self.peek().val in ("-", "+"):
                sgn = self.next().val
                ex = ("un", sgn, self.parse_postfix())
            else:
                ex = self.parse_postfix()
            base = ("bin", op, base, ex)
        return base

    def parse_postfix(self):
        e = self.parse_primary()
        while True:
            tok = self.peek()
            if tok.kind != "op":
                break
            if tok.val == "(" and not (self.in_matrix and self._paren_in_matrix == 0 and tok.sp):
                self.next()
                args = self.parse_args(")")
                e = ("index", e, args)
            elif tok.val == "." and self.peek(1).kind == "id" and not tok.sp:
                self.next()
                e = ("field", e, self.next().val)
            elif tok.val == "." and self.peek(1).val == "(" and not tok.sp:
                self.next()                      # dynamic field s.(name): parsed, not evaluated
                self.next()
                name = self.parse_expr()
                self.expect(")")
                e = ("dynfield", e, name)
            elif tok.val == "{" and not (self.in_matrix and tok.sp):
                self.next()                      # cell indexing c{...}: parsed, not evaluated
                args = self.parse_args("}")
                e = ("cellindex", e, args)
            elif tok.val == "@" and e[0] == "name" and self.peek(1).kind == "id" and not tok.sp:
                self.next()                      # superclass call  obj@Super(...) / method@Super(...)
                sup = self.next().val
                while self.at("."):
                    self.next()
                    sup += "." + self.next().val
                e = ("name", e[1] + "@" + sup)
            elif tok.val in ("'", ".'"):
                self.next()
                e = ("un", "'", e)
            else:
                break
        return e

    def parse_args(self, closer):
        args = []
        self.idx_depth += 1
        saved_matrix, saved_pim = self.in_matrix, self._paren_in_matrix
        if self.in_matrix:
            self._paren_in_matrix += 1
        while self.peek().kind == "nl":
            self.next()
        while not self.at(closer):
            if self.at(","):
                self.next()
                continue
            if self.at(":") and self.peek(1).val in (",", closer):
                self.next()
                args.append(("colon",))
                continue
            args.append(self.parse_expr())
            while self.peek().kind == "nl":
                self.next()
        self.next()
        self.idx_depth -= 1
        self.in_matrix, self._paren_in_matrix = saved_matrix, saved_pim
        return args

    def parse_primary(self):
        tok = self.next()
        if tok.kind == "num":
            return ("num", float(tok.val))
        if tok.kind == "str":
            return ("str", tok.val)
        if tok.kind == "id":
            if tok.val == "end" and self.idx_depth > 0:
                return ("end",)
            if tok.val in KEYWORDS:
                raise SyntaxError(f"{self.fname}:{tok.line}: unexpected keyword {tok.val!r}")
            return ("name", tok.val)
        if tok.val == "(":
            saved_matrix, saved_pim = self.in_matrix, self._paren_in_matrix
            if self.in_matrix:
                self._paren_in_matrix += 1
            saved_idx = self.idx_depth
            e = self.parse_expr()
            self.expect(")")
            self.in_matrix, self._paren_in_matrix = saved_matrix, saved_pim
            self.idx_depth = saved_idx
            return ("paren", e)
        if tok.val == "[":
            return self.parse_matrix()
        if tok.val == "@":
            if self.at("("):
                self.next()
                params = []
                while not self.at(")"):
                    if self.at(","):
                        self.next()
                        continue
                    params.append(self.next().val)
                self.next()
                body = self.parse_expr()
                return ("anon", params, body)
            name = self.next().val
            e = ("name", name)
            while self.at(".") and self.peek(1).kind == "id":
                self.next()
                e = ("field", e, self.next().val)
            return ("fhandle", e)
        if tok.val == ":":
            return ("colon",)
        raise SyntaxError(f"{self.fname}:{tok.line}: unexpected token {tok.val!r}")

    def parse_matrix(self):
        rows, row = [], []
        saved = (self.in_matrix, self._paren_in_matrix, self.idx_depth)
        self.in_matrix, self._paren_in_matrix = 1, 0
        idx_saved = self.idx_depth
        self.idx_depth = 0 if False else self.idx_depth   # `end` keeps its meaning of the enclosing index
        while True:
            tok = self.peek()
            if tok.kind == "op" and tok.val == "]":
                self.next()
                break
            if tok.kind == "nl" or (tok.kind == "op" and tok.val == ";"):
                self.next()
                if row:
                    rows.append(row)
                    row = []
                continue
            if tok.kind == "op" and tok.val == ",":
                self.next()
                continue
            row.append(self.parse_expr())
        if row:
            rows.append(row)
        self.in_matrix, self._paren_in_matrix, _ = saved
        self.idx_depth = idx_saved
        return ("matrix", rows)


# ------------------------------------------------------------------------------ runtime values
class MObject:
    """Instance of a handle classdef."""

    def __init__(self, cls):
        self.cls = cls
        self.props = {}

    def __repr__(self):
        return f"<{self.cls['name']} object>"


class MStruct(dict):
    pass


class MCell(list):
    """A MATLAB cell array (row)."""


class Closure:
    def __init__(self, interp, fdef, parent_frame=None, bound_obj=None, cls=None):
        self.interp, self.fdef, self.parent, self.bound, self.cls = interp, fdef, parent_frame, bound_obj, cls

    def __call__(self, args, nargout=1):
        if self.bound is not None:
            args = [self.bound] + list(args)
        return self.interp.call_function(self.fdef, args, nargout, self.parent, self.cls)


class Anon:
    def __init__(self, interp, params, body, frame):
        self.interp, self.params, self.body, self.frame = interp, params, body, dict(frame)

    def __call__(self, args, nargout=1):
        fr = Frame(self.frame)
        for p, a in zip(self.params, args):
            fr.vars[p] = a
        return [self.interp.eval(self.body, fr)]


class Frame:
    def __init__(self, parent_vars=None):
        self.vars = {} if parent_vars is None else dict(parent_vars)
        self.nargin = 0
        self.nargout = 0
        self.nested = {}
        self.parent = None


class _Return(Exception):
    pass


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class MatlabError(RuntimeError):
    pass


# ------------------------------------------------------------------------------ interpreter
class Interpreter:
    def __init__(self, path):
        """path: directories searched for `name.m` (classes, functions, scripts)."""
        self.path = list(path)
        self.cache = {}
        self.builtins = B.make_builtins(self)
        self.output = []

    # -- loading
    def find(self, name):
        if name in self.cache:
            return self.cache[name]
        for d in self.path:
            f = os.path.join(d, name + ".m")
            if os.path.isfile(f):
                with open(f, "r", encoding="utf-8", errors="replace") as fh:
                    src = fh.read()
                kind, node = Parser(tokenize(src), f).parse_file()
                self.cache[name] = (kind, node, f)
                return self.cache[name]
        self.cache[name] = None
        return None

    def run_script(self, name_or_src, frame=None, is_source=False):
        frame = frame or Frame()
        if is_source:
            kind, node = Parser(tokenize(name_or_src), "<script>").parse_file()
        else:
            ent = self.find(name_or_src)
            if ent is None:
                raise MatlabError(f"script {name_or_src} not found")
            kind, node, _ = ent
        if kind != "script":
            raise MatlabError("not a script")
        self.exec_block(node, frame)
        return frame

    # -- classes
    def class_chain(self, cls):
        chain = [cls]
        for s in cls["supers"]:
            if s == "handle":
                continue
            ent = self.find(s)
            if ent and ent[0] == "classdef":
                chain += self.class_chain(ent[1])
        return chain

    def find_method(self, cls, name):
        for c in self.class_chain(cls):
            if name in c["methods"]:
                return c["methods"][name], c
        return None, None

    def construct(self, cls, args):
        obj = MObject(cls)
        for c in reversed(self.class_chain(cls)):
            for pname, default in c["props"]:
                obj.props[pname] = B.empty() if default is None else self.eval(default, Frame())
        ctor = cls["methods"].get(cls["name"])
        if ctor is not None:
            fr = Frame()
            fr.nargin = len(args)
            fr.nargout = 1
            out = ctor["outs"][0]
            fr.vars[out] = obj
            cparams = ctor["params"]
            if cparams and cparams[-1] == "varargin":
                fr.vars["varargin"] = MCell(args[len(cparams) - 1:])
                args = args[:len(cparams) - 1]
                cparams = cparams[:-1]
            for p, a in zip(cparams, args):
                if p != "~":
                    fr.vars[p] = a
            fr.nested = ctor["nested"]
            try:
                self.exec_block(ctor["body"], fr)
            except _Return:
                pass
            obj = fr.vars[out]
        return obj

    def eval_args(self, arg_nodes, fr):
        """Evaluates call arguments; c{:} expands into several arguments."""
        out = []
        for a in arg_nodes:
            if a[0] == "cellindex" and len(a[2]) == 1 and a[2][0][0] == "colon":
                out.extend(self.eval(a[1], fr))
            else:
                out.append(self.eval(a, fr))
        return out

    def super_call(self, name, args, nargout, fr):
        """obj@Super(args...) inside a constructor, or method@Super(obj, args...)."""
        left, sup = name.split("@", 1)
        ent = self.find(sup)
        if ent is None or ent[0] != "classdef":
            raise MatlabError(f"unknown superclass {sup}")
        cls = ent[1]
        if left in fr.vars and isinstance(fr.vars[left], MObject):       # superclass constructor
            obj = fr.vars[left]
            ctor = cls["methods"].get(cls["name"])
            if ctor is not None:
                f2 = Frame()
                f2.vars[ctor["outs"][0]] = obj
                cparams = ctor["params"]
                if cparams and cparams[-1] == "varargin":
                    f2.vars["varargin"] = MCell(args[len(cparams) - 1:])
                    args = args[:len(cparams) - 1]
                    cparams = cparams[:-1]
                for p_, a_ in zip(cparams, args):
                    if p_ != "~":
                        f2.vars[p_] = a_
                f2.nargin, f2.nargout = len(args), 1
                f2.nested = ctor["nested"]
                try:
                    self.exec_block(ctor["body"], f2)
                except _Return:
                    pass
            return [obj]
        meth, owner = self.find_method(cls, left)                         # superclass method
        if meth is None:
            raise MatlabError(f"{sup} has no method {left}")
        return self.call_function(meth, args, nargout, None, owner)

    # -- calls
    def call_function(self, fdef, args, nargout, parent_frame=None, cls=None):
        fr = Frame()
        if parent_frame is not None:     # nested function: shares the parent's workspace
            fr.vars = parent_frame.vars
            fr.parent = parent_frame
        params = fdef["params"]
        if params and params[-1] == "varargin":
            nfix = len(params) - 1
            fr.vars["varargin"] = MCell(args[nfix:])
            args = args[:nfix]
            params = params[:nfix]
        if len(args) > len(params):
            raise MatlabError(f"too many input arguments to {fdef['name']}")
        saved = {}
        for p, a in zip(params, args):
            if p != "~":
                if parent_frame is not None and p in fr.vars:
                    saved[p] = fr.vars[p]
                fr.vars[p] = a
        fr.nargin, fr.nargout = len(args), nargout
        fr.nested = dict(fdef["nested"])
        if parent_frame is not None:
            fr.nested.update(parent_frame.nested)
        try:
            self.exec_block(fdef["body"], fr)
        except _Return:
            pass
        outs = []
        for k, o in enumerate(fdef["outs"][:max(nargout, 1)]):
            if o not in fr.vars:
                if k < nargout or (k == 0 and nargout <= 1 and False):
                    raise MatlabError(f"output argument {o!r} of {fdef['name']} not assigned")
                break
            outs.append(fr.vars[o])
        return outs

    def call_value(self, f, args, nargout):
        if isinstance(f, (Closure, Anon)):
            return f(args, nargout)
        if callable(f):
            r = f(*args, nargout=nargout) if getattr(f, "_wants_nargout", False) else f(*args)
            return list(r) if isinstance(r, tuple) else [r]
        raise MatlabError("value is not callable")

    def call_named(self, name, args, nargout, frame):
        if "@" in name:
            return self.super_call(name, args, nargout, frame)
        # nested / local functions
        fr = frame
        if name in frame.nested:
            return Closure(self, frame.nested[name], frame)(args, nargout)
        # method call in function syntax: method(obj, ...)
        if args and isinstance(args[0], MObject):
            meth, owner = self.find_method(args[0].cls, name)
            if meth is not None:
                return self.call_function(meth, args, nargout, None, owner)
        if args and hasattr(args[0], "call_method") and name in getattr(args[0], "METHODS", ()):
            return args[0].call_method(name, args[1:], nargout)
        if name in self.builtins:
            return self.call_value(self.builtins[name], args, nargout)
        ent = self.find(name)
        if ent is None:
            raise MatlabError(f"undefined function or variable {name!r}")
        kind, node, _ = ent
        if kind == "classdef":
            return [self.construct(node, args)]
        if kind == "functions":
            main = node[0]
            f = dict(main)
            f["nested"] = dict(main["nested"])
            for other in node[1:]:
                f["nested"][other["name"]] = other
            return self.call_function(f, args, nargout)
        # a script called by name
        self.exec_block(node, frame)
        return []

    # -- statements
    def exec_block(self, stmts, fr):
        for s in stmts:
            self.exec(s, fr)

    def exec(self, s, fr):
        k = s[0]
        if k == "assign":
            val = self.eval(s[2], fr, nargout=1)
            self.assign(s[1], val, fr)
        elif k == "massign":
            lhs = s[1]
            vals = self.eval_multi(s[2], fr, len(lhs))
            if len(vals) < len([x for x in lhs]):
                # fewer outputs than requested is an error only for the ones actually bound
                pass
            for target, v in zip(lhs, vals):
                if target is not None:
                    self.assign(target, v, fr)
            for target in lhs[len(vals):]:
                if target is not None:
                    raise MatlabError("too many output arguments requested")
        elif k == "expr":
            e = s[1]
            if e[0] == "name" and e[1] not in fr.vars:
                self.call_named(e[1], [], 0, fr)
            else:
                self.eval_multi(e, fr, 0)
        elif k == "if":
            for cond, body in s[1]:
                if B.truth(self.eval(cond, fr)):
                    self.exec_block(body, fr)
                    return
            if s[2] is not None:
                self.exec_block(s[2], fr)
        elif k == "for":
            rng = self.eval(s[2], fr)
            rng = B.as_array(rng)
            cols = rng.reshape(rng.shape[0], -1, order="F") if rng.ndim > 1 else rng.reshape(1, -1)
            for j in range(cols.shape[1]):
                col = cols[:, j]
                fr.vars[s[1]] = col.reshape(-1, 1).copy() if col.size > 1 else np.array([[col[0]]])
                try:
                    self.exec_block(s[3], fr)
                except _Break:
                    break
                except _Continue:
                    continue
        elif k == "while":
            while B.truth(self.eval(s[1], fr)):
                try:
                    self.exec_block(s[2], fr)
                except _Break:
                    break
                except _Continue:
                    continue
        elif k == "return":
            raise _Return()
        elif k == "break":
            raise _Break()
        elif k == "continue":
            raise _Continue()
        elif k == "nop":
            pass
        else:
            raise MatlabError(f"unknown statement {k}")

    # -- assignment
    def assign(self, target, val, fr):
        if isinstance(val, np.ndarray):
            val = val.copy()          # value semantics
        k = target[0]
        if k == "name":
            fr.vars[target[1]] = val
            return
        # walk the access chain down to the root variable
        chain = []
        node = target
        while node[0] in ("index", "field"):
            chain.append(node)
            node = node[1]
        if node[0] != "name":
            raise MatlabError("invalid assignment target")
        root = node[1]
        chain.reverse()
        cur = fr.vars.get(root)
        if cur is None:
            cur = MStruct() if chain[0][0] == "field" else B.empty()
            fr.vars[root] = cur
        new_root = self._assign_into(cur, chain, val, fr)
        fr.vars[root] = new_root

    def _assign_into(self, container, chain, val, fr):
        step = chain[0]
        last = len(chain) == 1
        if step[0] == "field":
            name = step[2]
            if isinstance(container, MObject):
                if last:
                    container.props[name] = val
                else:
                    container.props[name] = self._assign_into(container.props.get(name, B.empty()), chain[1:], val, fr)
                return container           # handle semantics
            if isinstance(container, np.ndarray) and container.size == 0:
                container = MStruct()
            if isinstance(container, MStruct):
                container = MStruct(container)
                if last:
                    container[name] = val
                else:
                    container[name] = self._assign_into(container.get(name, B.empty()), chain[1:], val, fr)
                return container
            raise MatlabError("field assignment into a non-struct")
        # indexed assignment
        if not last:
            raise MatlabError("indexing followed by further access on the left-hand side is not supported")
        arr = B.as_array(container).copy()
        idx = [self.eval_index(a, fr, arr, d, len(step[2])) for d, a in enumerate(step[2])]
        return B.index_assign(arr, idx, val)

    # -- expressions
    def eval_multi(self, e, fr, nargout):
        """Evaluates a call-like expression for `nargout` outputs; returns a list."""
        k = e[0]
        if k == "index":
            base = e[1]
            # obj.method(args) / struct.fhandle(args) / name(args)
            if base[0] == "name" and base[1] not in fr.vars:
                args = self.eval_args(e[2], fr)
                return self.call_named(base[1], args, nargout, fr)
            if base[0] == "field":
                owner = self.eval(base[1], fr)
                if isinstance(owner, MObject):
                    meth, cls = self.find_method(owner.cls, base[2])
                    if meth is not None and base[2] not in owner.props:
                        args = [owner] + self.eval_args(e[2], fr)
                        return self.call_function(meth, args, nargout, None, cls)
                if hasattr(owner, "call_method"):                  # built-in objects (inputParser)
                    return owner.call_method(base[2], self.eval_args(e[2], fr), nargout)
            target = self.eval(base, fr)
            if isinstance(target, (Closure, Anon)) or (callable(target) and not isinstance(target, np.ndarray)):
                args = self.eval_args(e[2], fr)
                return self.call_value(target, args, nargout)
            arr = B.as_array(target)
            idx = [self.eval_index(a, fr, arr, d, len(e[2])) for d, a in enumerate(e[2])]
            return [B.index(arr, idx)]
        if k == "name" and e[1] not in fr.vars:
            return self.call_named(e[1], [], nargout, fr)
        if k == "field":
            owner = self.eval(e[1], fr)
            if isinstance(owner, MObject) and e[2] not in owner.props:
                meth, cls = self.find_method(owner.cls, e[2])
                if meth is not None:
                    return self.call_function(meth, [owner], nargout, None, cls)
        return [self.eval(e, fr)]

    def eval_index(self, a, fr, arr, dim, ndims_used):
        if a[0] == "colon":
            return ("all",)
        end_val = B.end_value(arr, dim, ndims_used)
        fr2 = fr
        saved = fr.vars.get("__end__")
        fr.vars["__end__"] = end_val
        try:
            v = self.eval(a, fr2)
        finally:
            if saved is None:
                fr.vars.pop("__end__", None)
            else:
                fr.vars["__end__"] = saved
        return v

    def eval(self, e, fr, nargout=1):
        k = e[0]
        if k == "num":
            return np.array([[e[1]]])
        if k == "str":
            return e[1]
        if k == "name":
            name = e[1]
            if name in fr.vars:
                return fr.vars[name]
            if name == "nargin":
                return np.array([[float(fr.nargin)]])
            if name == "nargout":
                return np.array([[float(fr.nargout)]])
            r = self.call_named(name, [], 1, fr)
            if not r:
                raise MatlabError(f"{name} returned nothing")
            return r[0]
        if k == "end":
            return np.array([[float(fr.vars["__end__"])]])
        if k == "paren":
            return self.eval(e[1], fr)
        if k == "colon":
            return ":"
        if k == "bin":
            op = e[1]
            if op == "&&":
                return np.array([[float(B.truth(self.eval(e[2], fr)) and B.truth(self.eval(e[3], fr)))]])
            if op == "||":
                return np.array([[float(B.truth(self.eval(e[2], fr)) or B.truth(self.eval(e[3], fr)))]])
            return B.binop(op, self.eval(e[2], fr), self.eval(e[3], fr))
        if k == "un":
            return B.unop(e[1], self.eval(e[2], fr))
        if k == "range":
            a = self.eval(e[1], fr)
            b = self.eval(e[3], fr)
            s = self.eval(e[2], fr) if e[2] is not None else np.array([[1.0]])
            return B.colon(a, s, b)
        if k == "matrix":
            rows = []
            for row in e[1]:
                vals = [self.eval(x, fr) for x in row]
                rows.append(B.hcat(vals))
            return B.vcat(rows)
        if k == "field":
            owner = self.eval(e[1], fr)
            if isinstance(owner, MObject):
                if e[2] in owner.props:
                    return owner.props[e[2]]
                meth, cls = self.find_method(owner.cls, e[2])
                if meth is not None:
                    r = self.call_function(meth, [owner], 1, None, cls)
                    return r[0]
                raise MatlabError(f"no property or method {e[2]!r} on {owner!r}")
            if isinstance(owner, MStruct):
                if e[2] not in owner:
                    raise MatlabError(f"reference to non-existent field {e[2]!r}")
                return owner[e[2]]
            if hasattr(owner, "getfield"):
                return owner.getfield(e[2])
            if hasattr(owner, "call_method") and e[2] in getattr(owner, "METHODS", ()):
                return (owner.call_method(e[2], [], 1) or [B.empty()])[0]
            raise MatlabError(f"cannot take field {e[2]!r} of {type(owner).__name__}")
        if k == "index":
            r = self.eval_multi(e, fr, nargout)
            if not r:
                raise MatlabError("expression produced no value")
            return r[0]
        if k == "cellindex":
            c = self.eval(e[1], fr)
            if not isinstance(c, MCell):
                raise MatlabError("brace indexing into a non-cell")
            idx = self.eval(e[2][0], fr)
            return c[int(B.scalar(idx)) - 1]
        if k == "dynfield":
            owner = self.eval(e[1], fr)
            name = self.eval(e[2], fr)
            if isinstance(owner, MStruct):
                return owner[name]
            if isinstance(owner, MObject):
                return owner.props[name]
            raise MatlabError("dynamic field of a non-struct")
        if k == "anon":
            return Anon(self, e[1], e[2], fr.vars)
        if k == "fhandle":
            target = e[1]
            if target[0] == "name":
                name = target[1]
                if name in fr.nested:
                    return Closure(self, fr.nested[name], fr)
                return B.named_handle(self, name, fr)
            owner = self.eval(target[1], fr)
            meth, cls = self.find_method(owner.cls, target[2])
            return Closure(self, meth, None, owner, cls)
        raise MatlabError(f"cannot evaluate {k}")

"""mini_matlab -- a minimal MATLAB-subset interpreter (TEST INFRASTRUCTURE).

Purpose: neither MATLAB nor GNU Octave exists in this image, so the reference cannot be run
directly.  This module parses and executes the reference's *unmodified* ``.m`` files from
/root/reference (function files of the linear hot path: Normalize2Ddata, linearTFT,
transform_TFT, R_t_from_TFT, LinearTFTPoseEstimation, triangulation3D, ReprError, linearF,
LinearFPoseEstimation, TFT_from_P, crossM, AngError, project3Dpoints, generateSyntheticScene)
with NumPy/LAPACK standing in for the MATLAB runtime's built-ins (``svd``, ``inv``, ``rank``,
``kron``, ...).  Its outputs pin ``oracle/reference_port.py``: a transcription slip in the
restatement (a wrong sign in a design-matrix row, a transposed reshape) shows up as a mismatch
against the reference's own source text.  It is NOT a general MATLAB implementation: only the
syntax and built-ins those files use are supported, and it fails loudly on anything else.

Values: every numeric value is a NumPy array with ndim >= 2 (MATLAB semantics: scalars are 1x1,
column-major linear indexing / reshape); cell arrays are ``Cell`` (a list with a 2-D shape);
strings are ``str``; function handles are ``FuncHandle``.
"""
import os
import re

import numpy as np


class MatlabError(Exception):
    """Raised by the interpreted code's error(...)."""


class Cell(list):
    def __init__(self, items, shape=None):
        super().__init__(items)
        self.shape = shape if shape is not None else (1, len(items))


class FuncHandle:
    def __init__(self, name, scope=None):
        self.name = name
        self.scope = scope       # local functions of the file that created the handle (MATLAB keeps them reachable)


# ============================================================================ tokenizer
TOKEN_RE = re.compile(r"""
    (?P<num>(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?))
  | (?P<id>[A-Za-z_]\w*)
  | (?P<op>\.\*|\./|\.\^|\.\\|\.'|==|~=|<=|>=|&&|\|\||[-+*/\\^<>=&|~:(),;\[\]{}@'])
""", re.X)

KEYWORDS = {"function", "end", "if", "elseif", "else", "for", "while", "switch", "case", "otherwise",
            "return", "break", "continue"}


class Tok:
    __slots__ = ("kind", "val", "ws_before", "ws_after", "line")

    def __init__(self, kind, val, ws_before, line):
        self.kind, self.val, self.ws_before, self.ws_after, self.line = kind, val, ws_before, False, line

    def __repr__(self):
        return "%s:%r" % (self.kind, self.val)


def tokenize(src):
    toks = []
    i, n, line = 0, len(src), 1
    ws = True
    depth = 0          # bracket depth ([] and {}) for quote disambiguation
    while i < n:
        c = src[i]
        if c in " \t":
            i += 1; ws = True
            continue
        if c == "%":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if src.startswith("...", i):
            while i < n and src[i] != "\n":
                i += 1
            i += 1; line += 1; ws = True
            continue
        if c == "\n" or c == "\r":
            if c == "\n":
                toks.append(Tok("nl", "\n", ws, line)); line += 1
            i += 1; ws = True
            continue
        if c == "'":
            prev = toks[-1] if toks else None
            is_transpose = (prev is not None and not ws and
                            (prev.kind in ("num", "id") and prev.val not in KEYWORDS or prev.val in (")", "]", "}", "'", ".'")))
            if not is_transpose:
                j = i + 1; s = ""
                while j < n:
                    if src[j] == "'":
                        if j + 1 < n and src[j + 1] == "'":
                            s += "'"; j += 2; continue
                        break
                    s += src[j]; j += 1
                toks.append(Tok("str", s, ws, line)); i = j + 1; ws = False
                continue
        m = TOKEN_RE.match(src, i)
        if not m:
            raise SyntaxError("mini_matlab: cannot tokenize %r at line %d" % (src[i:i + 20], line))
        kind = m.lastgroup
        val = m.group(kind)
        if kind == "id" and val in KEYWORDS:
            kind = "kw"
        toks.append(Tok(kind, val, ws, line))
        i = m.end(); ws = False
    toks.append(Tok("nl", "\n", True, line))
    toks.append(Tok("eof", None, True, line))
    for a, b in zip(toks, toks[1:]):
        a.ws_after = b.ws_before
    return toks


# ============================================================================ parser -> AST (tuples)
class Parser:
    def __init__(self, toks, fname="?"):
        self.t, self.p, self.fname = toks, 0, fname

    def peek(self, k=0):
        return self.t[self.p + k]

    def next(self):
        tok = self.t[self.p]; self.p += 1
        return tok

    def at(self, val):
        return self.peek().val == val and self.peek().kind in ("op", "kw", "nl")

    def expect(self, val):
        tok = self.next()
        if tok.val != val:
            raise SyntaxError("%s:%d: expected %r, got %r" % (self.fname, tok.line, val, tok.val))
        return tok

    def skip_nl(self):
        while self.peek().kind == "nl" or self.at(";") or self.at(","):
            self.next()

    # ---- file ---------------------------------------------------------------------
    def parse_file(self):
        funcs = []
        self.skip_nl()
        while self.peek().kind != "eof":
            if not self.at("function"):
                raise SyntaxError("%s:%d: only function files are supported" % (self.fname, self.peek().line))
            funcs.append(self.parse_function())
            self.skip_nl()
        return funcs

    def parse_function(self):
        self.expect("function")
        outs = []
        if self.at("["):
            self.next()
            while not self.at("]"):
                if self.at(","):
                    self.next(); continue
                outs.append(self.next().val)
            self.next(); self.expect("=")
            name = self.next().val
        else:
            name = self.next().val
            if self.at("="):
                self.next(); outs = [name]; name = self.next().val
        params = []
        if self.at("("):
            self.next()
            while not self.at(")"):
                if self.at(","):
                    self.next(); continue
                params.append(self.next().val)
            self.next()
        body = self.parse_block(("end", "function"))
        if self.at("end"):
            self.next()
        return ("function", name, params, outs, body)

    def parse_block(self, stops):
        stmts = []
        while True:
            self.skip_nl()
            tok = self.peek()
            if tok.kind == "eof" or (tok.kind == "kw" and tok.val in stops):
                return stmts
            stmts.append(self.parse_statement())

    # ---- statements ---------------------------------------------------------------
    def parse_statement(self):
        tok = self.peek()
        if tok.kind == "kw":
            if tok.val == "if":
                return self.parse_if()
            if tok.val == "for":
                self.next()
                paren = self.at("(")
                if paren:
                    self.next()
                var = self.next().val; self.expect("=")
                rng = self.parse_expr()
                if paren:
                    self.expect(")")
                body = self.parse_block(("end",)); self.expect("end")
                return ("for", var, rng, body)
            if tok.val == "while":
                self.next(); cond = self.parse_expr()
                body = self.parse_block(("end",)); self.expect("end")
                return ("while", cond, body)
            if tok.val == "switch":
                self.next(); subj = self.parse_expr(); self.skip_nl()
                cases, default = [], None
                while not self.at("end"):
                    if self.at("case"):
                        self.next(); val = self.parse_expr()
                        cases.append((val, self.parse_block(("case", "otherwise", "end"))))
                    elif self.at("otherwise"):
                        self.next(); default = self.parse_block(("case", "otherwise", "end"))
                    else:
                        raise SyntaxError("%s:%d: bad switch" % (self.fname, self.peek().line))
                self.expect("end")
                return ("switch", subj, cases, default)
            if tok.val in ("return", "break", "continue"):
                self.next()
                return (tok.val,)
            raise SyntaxError("%s:%d: unexpected keyword %r" % (self.fname, tok.line, tok.val))
        # multi-assignment  [a,~,c] = f(...)
        if self.at("["):
            save = self.p
            targets = self.try_parse_lhs_list()
            if targets is not None:
                rhs = self.parse_expr()
                return ("multiassign", targets, rhs)
            self.p = save
        expr = self.parse_expr()
        if self.at("="):
            self.next()
            rhs = self.parse_expr()
            return ("assign", expr, rhs)
        return ("expr", expr)

    def try_parse_lhs_list(self):
        self.expect("[")
        targets = []
        while not self.at("]"):
            if self.at(","):
                self.next(); continue
            if self.at("~"):
                self.next(); targets.append(None); continue
            if self.peek().kind != "id":
                return None
            e = self.parse_postfix(False)
            targets.append(e)
        self.next()
        if not self.at("="):
            return None
        self.next()
        return targets

    def parse_if(self):
        self.expect("if")
        clauses = []
        cond = self.parse_expr()
        body = self.parse_block(("elseif", "else", "end"))
        clauses.append((cond, body))
        other = None
        while True:
            if self.at("elseif"):
                self.next(); c = self.parse_expr()
                clauses.append((c, self.parse_block(("elseif", "else", "end"))))
            elif self.at("else"):
                self.next(); other = self.parse_block(("end",))
            else:
                break
        self.expect("end")
        return ("if", clauses, other)

    # ---- expressions ----------------------------------------------------------------
    def parse_expr(self, inm=False):
        return self.parse_binary(0, inm)

    LEVELS = [("||",), ("&&",), ("|",), ("&",), ("==", "~=", "<", "<=", ">", ">="), (":",), ("+", "-"),
              ("*", "/", "\\", ".*", "./", ".\\")]

    def parse_binary(self, lvl, inm):
        if lvl == len(self.LEVELS):
            return self.parse_unary(inm)
        ops = self.LEVELS[lvl]
        if ops == (":",):
            return self.parse_range(lvl, inm)
        left = self.parse_binary(lvl + 1, inm)
        while True:
            tok = self.peek()
            if tok.kind != "op" or tok.val not in ops:
                return left
            if inm and tok.val in ("+", "-") and tok.ws_before and not tok.ws_after:
                return left                     # "[a -b]": two elements
            self.next()
            right = self.parse_binary(lvl + 1, inm)
            left = ("bin", tok.val, left, right)

    def parse_range(self, lvl, inm):
        first = self.parse_binary(lvl + 1, inm)
        if not (self.peek().kind == "op" and self.peek().val == ":") or self._colon_is_bare():
            return first
        self.next()
        second = self.parse_binary(lvl + 1, inm)
        if self.peek().kind == "op" and self.peek().val == ":" and not self._colon_is_bare():
            self.next()
            third = self.parse_binary(lvl + 1, inm)
            return ("range", first, second, third)
        return ("range", first, None, second)

    def _colon_is_bare(self):
        nxt = self.peek(1)
        return nxt.kind == "op" and nxt.val in (")", ",")

    def parse_unary(self, inm):
        tok = self.peek()
        if tok.kind == "op" and tok.val in ("-", "+", "~"):
            self.next()
            operand = self.parse_unary(inm)
            return ("un", tok.val, operand)
        return self.parse_power(inm)

    def parse_power(self, inm):
        base = self.parse_postfix(inm)
        while self.peek().kind == "op" and self.peek().val in ("^", ".^"):
            op = self.next().val
            if self.peek().kind == "op" and self.peek().val in ("-", "+", "~"):
                u = self.next().val
                exp = ("un", u, self.parse_postfix(inm))
            else:
                exp = self.parse_postfix(inm)
            base = ("bin", op, base, exp)
        return base

    def parse_postfix(self, inm):
        e = self.parse_primary(inm)
        while True:
            tok = self.peek()
            if tok.kind != "op":
                return e
            if tok.val == "(":
                if inm and tok.ws_before:
                    return e                    # "[a (b)]": two elements
                self.next()
                e = ("index", e, self.parse_args(")"))
            elif tok.val == "{":
                if inm and tok.ws_before:
                    return e
                self.next()
                e = ("cellindex", e, self.parse_args("}"))
            elif tok.val in ("'", ".'"):
                self.next()
                e = ("transpose", e)
            else:
                return e

    def parse_args(self, close):
        args = []
        while not self.at(close):
            if self.at(","):
                self.next(); continue
            if self.at(":") and self.peek(1).kind == "op" and self.peek(1).val in (",", close):
                self.next(); args.append(("colon",)); continue
            args.append(self.parse_expr())
        self.next()
        return args

    def parse_primary(self, inm):
        tok = self.next()
        if tok.kind == "num":
            return ("num", float(tok.val))
        if tok.kind == "str":
            return ("str", tok.val)
        if tok.kind == "id":
            return ("id", tok.val)
        if tok.kind == "kw" and tok.val == "end":
            return ("endkw",)
        if tok.kind == "op":
            if tok.val == "(":
                e = self.parse_expr(False)
                self.expect(")")
                return ("paren", e)
            if tok.val == "[":
                return self.parse_matrix("]")
            if tok.val == "{":
                m = self.parse_matrix("}")
                return ("cell", m[1])
            if tok.val == "@":
                return ("handle", self.next().val)
            if tok.val == ":":
                return ("colon",)
        raise SyntaxError("%s:%d: unexpected token %r" % (self.fname, tok.line, tok.val))

    def parse_matrix(self, close):
        rows, row = [], []
        while True:
            tok = self.peek()
            if tok.kind == "op" and tok.val == close:
                self.next()
                if row:
                    rows.append(row)
                return ("matrix", rows)
            if (tok.kind == "op" and tok.val == ";") or tok.kind == "nl":
                self.next()
                if row:
                    rows.append(row); row = []
                continue
            if tok.kind == "op" and tok.val == ",":
                self.next(); continue
            row.append(self.parse_expr(True))


# ============================================================================ runtime helpers
def _num(x):
    a = np.asarray(x)
    if a.dtype == bool:
        pass
    elif not np.iscomplexobj(a):
        a = a.astype(np.float64, copy=False)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(1, -1)
    return a


def _scalar(x):
    a = np.asarray(x)
    if a.size != 1:
        raise ValueError("mini_matlab: expected a scalar, got shape %s" % (a.shape,))
    v = a.reshape(-1)[0]
    return complex(v) if np.iscomplexobj(a) and v.imag != 0 else float(v.real if np.iscomplexobj(a) else v)


def _is_scalar(a):
    return isinstance(a, np.ndarray) and a.size == 1


def _truth(v):
    a = np.asarray(v)
    return a.size > 0 and bool(np.all(a != 0))


def _colvec(a):
    return a.reshape(-1, 1, order="F")


def _idx(v):
    """1-based MATLAB index array -> 0-based int array (column-major flattened) and its shape."""
    a = np.asarray(v)
    if a.dtype == bool:
        return np.flatnonzero(a.reshape(-1, order="F")), None
    return (np.rint(a.reshape(-1, order="F")).astype(np.int64) - 1), a.shape


class _Return(Exception):
    pass


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


# ============================================================================ interpreter
class Interpreter:
    """Loads function files from `paths` on demand and evaluates calls."""

    def __init__(self, paths, rng_factory=None):
        self.paths = list(paths)
        self.files = {}          # name -> (main function, {local name: function})
        self.rng_factory = rng_factory
        self.rng = None
        self.call_depth = 0

    # ---- function lookup ------------------------------------------------------------
    def load(self, name):
        if name in self.files:
            return self.files[name]
        for d in self.paths:
            f = os.path.join(d, name + ".m")
            if os.path.exists(f):
                funcs = Parser(tokenize(open(f).read()), f).parse_file()
                entry = (funcs[0], {fn[1]: fn for fn in funcs[1:]})
                self.files[name] = entry
                return entry
        return None

    def call(self, name, args, nargout=1, local_scope=None):
        """Call an .m function (or builtin) by name with already-evaluated args -> list of outputs."""
        fn = None
        locals_ = local_scope or {}
        if name in locals_:
            fn = locals_[name]
        else:
            entry = self.load(name)
            if entry is not None:
                fn, locals_ = entry
        if fn is None:
            return self.builtin(name, args, nargout)
        _, fname, params, outs, body = fn
        if len(args) > len(params):
            raise MatlabError("Too many input arguments to %s" % fname)
        env = {"__nargin__": len(args), "__nargout__": nargout, "__locals__": locals_}
        for p, a in zip(params, args):
            env[p] = a
        try:
            self.exec_block(body, env)
        except _Return:
            pass
        res = []
        for k, o in enumerate(outs[:max(1, nargout)]):
            if o not in env:
                if k < nargout:
                    raise MatlabError("Output argument '%s' (and maybe others) not assigned during call to '%s'" % (o, fname))
                break
            res.append(env[o])
        return res

    # ---- statements -------------------------------------------------------------------
    def exec_block(self, stmts, env):
        for s in stmts:
            self.exec_stmt(s, env)

    def exec_stmt(self, s, env):
        kind = s[0]
        if kind == "assign":
            self.assign(s[1], self.eval(s[2], env), env)
        elif kind == "multiassign":
            targets, rhs = s[1], s[2]
            vals = self.eval_multi(rhs, env, len(targets))
            for tgt, v in zip(targets, vals):
                if tgt is not None:
                    self.assign(tgt, v, env)
        elif kind == "expr":
            e = s[1]
            if e[0] == "id" and e[1] not in env:          # command-style call, e.g. "rng(seed)" parsed as index
                self.eval_multi(e, env, 0)
            else:
                self.eval_multi(e, env, 0)
        elif kind == "if":
            for cond, body in s[1]:
                if _truth(self.eval(cond, env)):
                    self.exec_block(body, env)
                    return
            if s[2] is not None:
                self.exec_block(s[2], env)
        elif kind == "for":
            rng = self.eval(s[2], env)
            cols = rng.reshape(rng.shape[0], -1, order="F") if isinstance(rng, np.ndarray) else rng
            ncols = cols.shape[1] if isinstance(cols, np.ndarray) else len(cols)
            for c in range(ncols):
                env[s[1]] = cols[:, c:c + 1].copy() if isinstance(cols, np.ndarray) else cols[c]
                try:
                    self.exec_block(s[3], env)
                except _Break:
                    break
                except _Continue:
                    continue
        elif kind == "while":
            while _truth(self.eval(s[1], env)):
                try:
                    self.exec_block(s[2], env)
                except _Break:
                    break
                except _Continue:
                    continue
        elif kind == "switch":
            subj = self.eval(s[1], env)
            for val, body in s[2]:
                v = self.eval(val, env)
                hit = (subj == v) if isinstance(subj, str) or isinstance(v, str) else (_scalar(subj) == _scalar(v))
                if hit:
                    self.exec_block(body, env)
                    return
            if s[3] is not None:
                self.exec_block(s[3], env)
        elif kind == "return":
            raise _Return()
        elif kind == "break":
            raise _Break()
        elif kind == "continue":
            raise _Continue()
        else:
            raise NotImplementedError(kind)

    def assign(self, target, value, env):
        if target[0] == "id":
            env[target[1]] = value.copy() if isinstance(value, np.ndarray) else value
            return
        if target[0] == "index" and target[1][0] == "id":
            name = target[1][1]
            if name not in env:
                raise MatlabError("indexed assignment to undefined variable %s" % name)
            A = env[name]
            args = target[2]
            subs = [self.eval_index_arg(a, env, A, k, len(args)) for k, a in enumerate(args)]
            val = _num(value)
            if len(subs) == 1:
                flat = A.reshape(-1, order="F").copy()
                ii = np.arange(flat.size) if subs[0] is None else subs[0]
                flat[ii] = val.reshape(-1, order="F") if val.size > 1 else val.reshape(-1)[0]
                env[name] = flat.reshape(A.shape, order="F")
            else:
                A = A.copy()
                while A.ndim < len(subs):
                    A = A[..., None]
                ix = [np.arange(A.shape[k]) if sidx is None else sidx for k, sidx in enumerate(subs)]
                if any((i.size and i.max() >= A.shape[k]) for k, i in enumerate(ix)):
                    raise MatlabError("mini_matlab: growing assignment not supported (%s)" % name)
                block_shape = tuple(i.size for i in ix)
                if val.size == 1:
                    A[np.ix_(*ix)] = val.reshape(-1)[0]
                else:
                    A[np.ix_(*ix)] = val.reshape(block_shape, order="F") if val.shape != block_shape else val
                env[name] = A
            return
        raise NotImplementedError("assignment target %r" % (target[0],))

    # ---- expressions ----------------------------------------------------------------------
    def eval_index_arg(self, a, env, A, k, nargs, want_shape=False):
        """Returns None for ':' else the 0-based index vector (and, on request, the index's own shape)."""
        if a[0] == "colon":
            return (None, None) if want_shape else None
        if isinstance(A, Cell):
            end_val = A.shape[k] if nargs > 1 else len(A)
        elif nargs == 1:
            end_val = A.size
        elif k == nargs - 1:
            end_val = int(np.prod(A.shape[k:])) if k < A.ndim else 1
        else:
            end_val = A.shape[k] if k < A.ndim else 1
        env_end = env.get("__end__")
        env["__end__"] = float(end_val)
        try:
            v = self.eval(a, env)
        finally:
            if env_end is None:
                env.pop("__end__", None)
            else:
                env["__end__"] = env_end
        if isinstance(v, str) and v == ":":
            return (None, None) if want_shape else None
        idx, shp = _idx(v)
        return (idx, shp) if want_shape else idx

    def eval(self, e, env):
        r = self.eval_multi(e, env, 1)
        if not r:
            raise MatlabError("expression produced no value")
        return r[0]

    def eval_multi(self, e, env, nargout):
        kind = e[0]
        if kind == "num":
            return [np.array([[e[1]]])]
        if kind == "str":
            return [e[1]]
        if kind == "colon":
            return [":"]
        if kind == "endkw":
            return [np.array([[env["__end__"]]])]
        if kind == "paren":
            return [self.eval(e[1], env)]
        if kind == "handle":
            return [FuncHandle(e[1], env.get("__locals__"))]
        if kind == "id":
            name = e[1]
            if name in env:
                return [env[name]]
            if name == "nargin":
                return [np.array([[float(env["__nargin__"])]])]
            if name == "nargout":
                return [np.array([[float(env["__nargout__"])]])]
            return self.call(name, [], nargout, env.get("__locals__"))
        if kind == "un":
            v = _num(self.eval(e[2], env))
            if e[1] == "-":
                return [-v]
            if e[1] == "+":
                return [v]
            return [np.logical_not(v != 0)]
        if kind == "transpose":
            v = self.eval(e[1], env)
            if isinstance(v, Cell):
                return [Cell(list(v), (v.shape[1], v.shape[0]))]
            v = _num(v)
            if v.ndim != 2:
                raise MatlabError("transpose on N-D array")
            return [v.T.copy()]
        if kind == "bin":
            return [self.binop(e[1], e[2], e[3], env)]
        if kind == "range":
            a = _scalar(self.eval(e[1], env)); b = _scalar(self.eval(e[3], env))
            step = 1.0 if e[2] is None else _scalar(self.eval(e[2], env))
            nsteps = int(np.floor((b - a) / step + 1e-10)) + 1 if (b - a) / step >= -1e-10 else 0
            return [(a + step * np.arange(max(nsteps, 0))).reshape(1, -1)]
        if kind == "matrix":
            return [self.build_matrix(e[1], env)]
        if kind == "cell":
            items = []
            for row in e[1]:
                items.extend(self.eval(x, env) for x in row)
            return [Cell(items)]
        if kind == "cellindex":
            c = self.eval(e[1], env)
            if not isinstance(c, Cell):
                raise MatlabError("brace indexing on a non-cell")
            subs = [self.eval_index_arg(a, env, c, k, len(e[2])) for k, a in enumerate(e[2])]
            if len(subs) == 1:
                return [c[int(subs[0][0])]]
            r, cc = int(subs[0][0]), int(subs[1][0])
            return [c[r + c.shape[0] * cc]]
        if kind == "index":
            base = e[1]
            if base[0] == "id" and base[1] not in env:
                args = [self.eval(a, env) for a in e[2]]
                return self.call(base[1], args, nargout, env.get("__locals__"))
            A = self.eval(base, env)
            if isinstance(A, FuncHandle):
                args = [self.eval(a, env) for a in e[2]]
                return self.call(A.name, args, nargout, A.scope if A.scope is not None else env.get("__locals__"))
            if isinstance(A, Cell):
                subs = [self.eval_index_arg(a, env, A, k, len(e[2])) for k, a in enumerate(e[2])]
                if len(subs) == 1:
                    ii = range(len(A)) if subs[0] is None else subs[0]
                    return [Cell([A[int(i)] for i in ii])]
                raise NotImplementedError("2-D paren indexing of cells")
            return [self.index(_num(A), e[2], env)]
        raise NotImplementedError(kind)

    def index(self, A, args, env):
        if len(args) == 1:
            idx, shp = self.eval_index_arg(args[0], env, A, 0, 1, want_shape=True)
            flat = A.reshape(-1, order="F")
            if idx is None:
                return flat.reshape(-1, 1).copy()
            out = flat[idx]
            # a vector source keeps its orientation, otherwise the result takes the index's shape
            if A.ndim == 2 and A.shape[0] == 1:
                return out.reshape(1, -1)
            if A.ndim == 2 and A.shape[1] == 1:
                return out.reshape(-1, 1)
            if shp is not None and len(shp) == 2:
                return out.reshape(shp, order="F")
            return out.reshape(-1, 1)
        subs = [self.eval_index_arg(a, env, A, k, len(args)) for k, a in enumerate(args)]
        if len(subs) < A.ndim:                     # fewer subscripts than dimensions: trailing ones collapse
            A = A.reshape(A.shape[:len(subs) - 1] + (-1,), order="F")
        while A.ndim < len(subs):
            A = A[..., None]
        ix = [np.arange(A.shape[k]) if sidx is None else sidx for k, sidx in enumerate(subs)]
        out = A[np.ix_(*ix)]
        while out.ndim > 2 and out.shape[-1] == 1:
            out = out[..., 0]
        return out.copy()

    def build_matrix(self, rows, env):
        if not rows:
            return np.zeros((0, 0))
        built = []
        for row in rows:
            parts = []
            for x in row:
                v = self.eval(x, env)
                if isinstance(v, str):
                    raise NotImplementedError("char arrays in matrix literals")
                v = _num(v)
                if v.size or v.shape[0]:
                    parts.append(v)
            parts = [p for p in parts if p.size > 0]
            if parts:
                built.append(np.concatenate(parts, axis=1) if len(parts) > 1 else parts[0])
        if not built:
            return np.zeros((0, 0))
        return np.concatenate(built, axis=0) if len(built) > 1 else built[0].copy()

    def binop(self, op, le, re_, env):
        if op == "&&":
            return np.array([[_truth(self.eval(le, env)) and _truth(self.eval(re_, env))]])
        if op == "||":
            return np.array([[_truth(self.eval(le, env)) or _truth(self.eval(re_, env))]])
        a = self.eval(le, env); b = self.eval(re_, env)
        if isinstance(a, str) or isinstance(b, str):
            if op == "==":
                return np.array([[a == b]])
            raise NotImplementedError("string operator %s" % op)
        a = _num(a); b = _num(b)
        if a.dtype == bool and op in ("+", "-", "*", ".*", "/", "./"):
            a = a.astype(np.float64)
        if b.dtype == bool and op in ("+", "-", "*", ".*", "/", "./"):
            b = b.astype(np.float64)
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == ".*":
            return a * b
        if op == "./":
            with np.errstate(divide="ignore", invalid="ignore"):
                return a / b
        if op == "*":
            if a.size == 1 or b.size == 1:
                return a * b
            if a.ndim != 2 or b.ndim != 2:
                raise MatlabError("matrix product of N-D arrays")
            return a @ b
        if op == "/":
            if b.size == 1:
                with np.errstate(divide="ignore", invalid="ignore"):
                    return a / b
            return np.linalg.solve(b.T, a.T).T
        if op == "\\":
            if a.size == 1:
                return b / a
            return np.linalg.solve(a, b) if a.shape[0] == a.shape[1] else np.linalg.lstsq(a, b, rcond=None)[0]
        if op == ".^":
            return np.power(a, b)
        if op == "^":
            if a.size == 1 and b.size == 1:
                return np.power(a, b)
            if b.size == 1 and float(b.reshape(-1)[0]).is_integer():
                return np.linalg.matrix_power(a, int(b.reshape(-1)[0]))
            raise NotImplementedError("matrix power")
        if op in ("==", "~=", "<", "<=", ">", ">="):
            f = {"==": np.equal, "~=": np.not_equal, "<": np.less, "<=": np.less_equal, ">": np.greater,
                 ">=": np.greater_equal}[op]
            return f(a, b)
        if op == "&":
            return np.logical_and(a != 0, b != 0)
        if op == "|":
            return np.logical_or(a != 0, b != 0)
        raise NotImplementedError(op)

    # ---- built-ins (NumPy/LAPACK standing in for the MATLAB runtime) ------------------------------
    def builtin(self, name, args, nargout):
        f = getattr(self, "bi_" + name, None)
        if f is None:
            raise NotImplementedError("mini_matlab: function or variable '%s' is not defined" % name)
        out = f(args, nargout)
        return out if isinstance(out, list) else [out]

    @staticmethod
    def _dimarg(args, k, default=None):
        return int(_scalar(args[k])) if len(args) > k else default

    @staticmethod
    def _first_nonsingleton(a):
        for d, s in enumerate(a.shape):
            if s != 1:
                return d
        return 0

    def bi_pi(self, a, n):
        return np.array([[np.pi]])

    def bi_eps(self, a, n):
        return np.array([[np.finfo(float).eps]]) if not a else np.spacing(np.abs(_num(a[0])))

    def bi_size(self, a, n):
        v = a[0]
        shp = v.shape if isinstance(v, (np.ndarray, Cell)) else ((1, len(v)) if isinstance(v, str) else (1, 1))
        if len(a) > 1:
            d = int(_scalar(a[1])) - 1
            return np.array([[float(shp[d] if d < len(shp) else 1)]])
        if n <= 1:
            return np.array([[float(s) for s in shp]])
        return [np.array([[float(shp[k] if k < len(shp) else 1)]]) for k in range(n)]

    def bi_length(self, a, n):
        v = a[0]
        if isinstance(v, Cell):
            return np.array([[float(len(v))]])
        v = _num(v)
        return np.array([[float(0 if v.size == 0 else max(v.shape))]])

    def bi_numel(self, a, n):
        return np.array([[float(len(a[0]) if isinstance(a[0], Cell) else _num(a[0]).size)]])

    def bi_isempty(self, a, n):
        v = a[0]
        return np.array([[(len(v) == 0) if isinstance(v, (Cell, str)) else (_num(v).size == 0)]])

    def _shape_args(self, a):
        if not a:
            return (1, 1)
        if len(a) == 1:
            v = _num(a[0])
            if v.size == 1:
                k = int(_scalar(v)); return (k, k)
            return tuple(int(x) for x in v.reshape(-1))
        return tuple(int(_scalar(x)) for x in a)

    def bi_zeros(self, a, n):
        return np.zeros(self._shape_args(a))

    def bi_ones(self, a, n):
        return np.ones(self._shape_args(a))

    def bi_eye(self, a, n):
        s = self._shape_args(a)
        return np.eye(s[0], s[1])

    def bi_repmat(self, a, n):
        reps = self._shape_args(a[1:])
        return np.tile(_num(a[0]), reps)

    def bi_reshape(self, a, n):
        v = _num(a[0])
        dims = a[1:]
        if len(dims) > 1 and any(_num(d).size == 0 for d in dims):        # reshape(x, 6, []): one dimension inferred
            known = [None if _num(d).size == 0 else int(_scalar(d)) for d in dims]
            prod = int(np.prod([k for k in known if k is not None])) if any(k is not None for k in known) else 1
            shp = tuple(v.size // prod if k is None else k for k in known)
        else:
            shp = self._shape_args(dims)
        return v.reshape(shp, order="F").copy()

    def bi_diag(self, a, n):
        v = _num(a[0])
        if 1 in v.shape:
            return np.diag(v.reshape(-1))
        return np.diag(v).reshape(-1, 1)

    def _reduce(self, fn, a):
        v = _num(a[0])
        if v.dtype == bool:
            v = v.astype(np.float64)
        d = (int(_scalar(a[1])) - 1) if len(a) > 1 else self._first_nonsingleton(v)
        return fn(v, axis=d, keepdims=True) if d < v.ndim else v

    def bi_sum(self, a, n):
        return self._reduce(np.sum, a)

    def bi_mean(self, a, n):
        return self._reduce(np.mean, a)

    def bi_min(self, a, n):
        if len(a) == 2:
            return np.minimum(_num(a[0]), _num(a[1]))
        return self._reduce(np.min, a)

    def bi_max(self, a, n):
        if len(a) == 2:
            return np.maximum(_num(a[0]), _num(a[1]))
        return self._reduce(np.max, a)

    def bi_sqrt(self, a, n):
        v = _num(a[0])
        return np.sqrt(v) if np.all(v >= 0) else np.sqrt(v.astype(complex))

    def bi_abs(self, a, n):
        return np.abs(_num(a[0]))

    def bi_sign(self, a, n):
        return np.sign(_num(a[0]))

    def bi_sin(self, a, n):
        return np.sin(_num(a[0]))

    def bi_cos(self, a, n):
        return np.cos(_num(a[0]))

    def bi_acos(self, a, n):
        v = _num(a[0])
        if np.all(np.abs(v) <= 1):
            return np.arccos(v)
        return np.arccos(v.astype(complex))        # MATLAB returns a complex angle for |x| > 1

    def bi_trace(self, a, n):
        return np.array([[np.trace(_num(a[0]))]])

    def bi_det(self, a, n):
        return np.array([[np.linalg.det(_num(a[0]))]])

    def bi_inv(self, a, n):
        return np.linalg.inv(_num(a[0]))

    def bi_pinv(self, a, n):
        """MATLAB pinv: SVD with tolerance max(size(A)) * eps(norm(A)); singular values <= tol are dropped."""
        v = _num(a[0])
        if v.size == 0:
            return np.zeros((v.shape[1], v.shape[0]))
        U, s, Vh = np.linalg.svd(v, full_matrices=False)
        tol = max(v.shape) * np.spacing(s.max()) if len(a) < 2 else _scalar(a[1])
        r = int(np.sum(s > tol))
        return (Vh[:r].T / s[:r]) @ U[:, :r].T

    def bi_any(self, a, n):
        v = _num(a[0])
        if v.ndim == 2 and 1 not in v.shape and v.size:
            return np.any(v != 0, axis=0, keepdims=True).astype(np.float64)
        return np.array([[float(np.any(v != 0))]])

    def bi_isnan(self, a, n):
        return np.isnan(_num(a[0])).astype(np.float64)

    def bi_isinf(self, a, n):
        return np.isinf(_num(a[0])).astype(np.float64)

    def bi_kron(self, a, n):
        return np.kron(_num(a[0]), _num(a[1]))

    def bi_norm(self, a, n):
        v = _num(a[0])
        if 1 in v.shape:
            return np.array([[np.linalg.norm(v.reshape(-1))]])
        return np.array([[np.linalg.norm(v, 2)]])

    def bi_rank(self, a, n):
        v = _num(a[0])
        s = np.linalg.svd(v, compute_uv=False)
        tol = max(v.shape) * np.spacing(s.max()) if s.size else 0.0
        return np.array([[float(np.sum(s > tol))]])

    def bi_svd(self, a, n):
        v = _num(a[0])
        if n <= 1:
            return np.linalg.svd(v, compute_uv=False).reshape(-1, 1)
        U, s, Vh = np.linalg.svd(v, full_matrices=True)
        S = np.zeros(v.shape)
        S[:s.size, :s.size] = np.diag(s)
        return [U, S, Vh.T.copy()]

    def bi_dot(self, a, n):
        x, y = _num(a[0]), _num(a[1])
        if len(a) > 2:
            return np.sum(x * y, axis=int(_scalar(a[2])) - 1, keepdims=True)
        if 1 in x.shape:
            return np.array([[np.dot(x.reshape(-1), y.reshape(-1))]])
        return np.sum(x * y, axis=0, keepdims=True)

    def bi_cross(self, a, n):
        x, y = _num(a[0]), _num(a[1])
        if len(a) > 2:
            return np.cross(x, y, axis=int(_scalar(a[2])) - 1)
        if x.size == 3:
            return np.cross(x.reshape(-1), y.reshape(-1)).reshape(x.shape)
        ax = 0 if x.shape[0] == 3 else 1
        return np.cross(x, y, axis=ax)

    def bi_find(self, a, n):
        v = _num(a[0])
        idx = np.flatnonzero(v.reshape(-1, order="F")).astype(np.float64) + 1
        return idx.reshape(1, -1) if v.shape[0] == 1 else idx.reshape(-1, 1)

    def bi_cell2mat(self, a, n):
        c = a[0]
        if c.shape[1] == 1:
            return np.concatenate([_num(x) for x in c], axis=0)
        if c.shape[0] == 1:
            return np.concatenate([_num(x) for x in c], axis=1)
        raise NotImplementedError("2-D cell2mat")

    def bi_bsxfun(self, a, n):
        op = {"rdivide": "./", "times": ".*", "plus": "+", "minus": "-"}[a[0].name]
        x, y = _num(a[1]), _num(a[2])
        with np.errstate(divide="ignore", invalid="ignore"):
            return {"./": x / y, ".*": x * y, "+": x + y, "-": x - y}[op]

    def bi_error(self, a, n):
        raise MatlabError(a[0] if a else "error")

    def bi_fprintf(self, a, n):
        return []

    def bi_disp(self, a, n):
        return []

    # RNG hooks: the documented substitute for MATLAB's generators (oracle/scene.py docstring)
    def bi_rng(self, a, n):
        if self.rng_factory is None:
            raise MatlabError("rng() needs an rng_factory")
        self.rng = self.rng_factory(int(_scalar(a[0])))
        return []

    def bi_rand(self, a, n):
        r, c = self._shape_args(a)
        return self.rng.rand(r, c)

    def bi_randn(self, a, n):
        r, c = self._shape_args(a)
        return self.rng.randn(r, c)

    def bi_randsample(self, a, n):
        pop = _num(a[0])
        k = int(_scalar(a[1]))
        if pop.size == 1:
            return (self.rng.randsample(int(_scalar(pop)), k).astype(np.float64) + 1).reshape(1, -1)
        flat = pop.reshape(-1, order="F")
        return flat[self.rng.randsample(flat.size, k)].reshape(1, -1)


def reference_interpreter(root="/root/reference", rng_factory=None):
    """Interpreter whose search path is the reference's function directories."""
    paths = [os.path.join(root, d) for d in ("TFT_methods", "F_methods", "auxiliar_functions", "Optimization", "Data")]
    return Interpreter(paths, rng_factory)

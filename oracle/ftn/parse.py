"""Parser for the Fortran subset the reference is written in (free form, modules, derived types with type-bound procedures,
allocatable arrays, array sections, internal procedures, list-directed / formatted / stream I/O).  TEST INFRASTRUCTURE, see
oracle/ftn/README.md.  Produces plain Python objects (class Node with a `t` tag)."""
from __future__ import annotations

from typing import List, Optional

from .lex import Tok, logical_lines, tokenize


class Node:
    def __init__(self, t, **kw):
        self.t = t
        self.__dict__.update(kw)

    def __repr__(self):
        return "Node(" + ", ".join(f"{k}={v!r}" for k, v in self.__dict__.items()) + ")"


TYPE_WORDS = ("integer", "real", "double", "logical", "character", "type", "class", "complex")


class ParseError(Exception):
    pass


# ---------------------------------------------------------------------------------------------------------
class ExprParser:
    def __init__(self, toks: List[Tok], pos: int = 0, where: str = ""):
        self.t, self.p, self.where = toks, pos, where

    # -- token helpers
    def peek(self, k=0) -> Optional[Tok]:
        return self.t[self.p + k] if self.p + k < len(self.t) else None

    def at_end(self):
        return self.p >= len(self.t)

    def is_op(self, v, k=0):
        tk = self.peek(k)
        return tk is not None and tk.kind == "op" and tk.val == v

    def is_id(self, v=None, k=0):
        tk = self.peek(k)
        return tk is not None and tk.kind == "id" and (v is None or tk.val == v)

    def eat_op(self, v):
        if not self.is_op(v):
            raise ParseError(f"expected {v!r} at token {self.p} in: {self.where} (got {self.peek()})")
        self.p += 1

    def eat_id(self, v=None):
        if not self.is_id(v):
            raise ParseError(f"expected identifier {v or ''} at token {self.p} in: {self.where} (got {self.peek()})")
        self.p += 1
        return self.t[self.p - 1].val

    # -- expressions, lowest precedence first
    def expr(self):
        return self.p_eqv()

    def p_eqv(self):
        l = self.p_or()
        while self.peek() and self.peek().kind == "dotop" and self.peek().val in ("eqv", "neqv"):
            op = self.peek().val; self.p += 1
            l = Node("bin", op=op, l=l, r=self.p_or())
        return l

    def p_or(self):
        l = self.p_and()
        while self.peek() and self.peek().kind == "dotop" and self.peek().val == "or":
            self.p += 1
            l = Node("bin", op="or", l=l, r=self.p_and())
        return l

    def p_and(self):
        l = self.p_not()
        while self.peek() and self.peek().kind == "dotop" and self.peek().val == "and":
            self.p += 1
            l = Node("bin", op="and", l=l, r=self.p_not())
        return l

    def p_not(self):
        if self.peek() and self.peek().kind == "dotop" and self.peek().val == "not":
            self.p += 1
            return Node("un", op="not", e=self.p_not())
        return self.p_rel()

    REL = {"==": "eq", "/=": "ne", "<": "lt", "<=": "le", ">": "gt", ">=": "ge"}

    def p_rel(self):
        l = self.p_concat()
        tk = self.peek()
        if tk is not None:
            op = None
            if tk.kind == "dotop" and tk.val in ("eq", "ne", "lt", "le", "gt", "ge"):
                op = tk.val
            elif tk.kind == "op" and tk.val in self.REL:
                op = self.REL[tk.val]
            if op:
                self.p += 1
                return Node("bin", op=op, l=l, r=self.p_concat())
        return l

    def p_concat(self):
        l = self.p_add()
        while self.is_op("//"):
            self.p += 1
            l = Node("bin", op="//", l=l, r=self.p_add())
        return l

    def p_add(self):
        if self.is_op("-") or self.is_op("+"):
            op = self.peek().val; self.p += 1
            l = self.p_mul()
            if op == "-":
                l = Node("un", op="neg", e=l)
        else:
            l = self.p_mul()
        while self.is_op("+") or self.is_op("-"):
            op = self.peek().val; self.p += 1
            l = Node("bin", op=op, l=l, r=self.p_mul())
        return l

    def p_mul(self):
        l = self.p_pow()
        while self.is_op("*") or self.is_op("/"):
            op = self.peek().val; self.p += 1
            l = Node("bin", op=op, l=l, r=self.p_pow())
        return l

    def p_pow(self):
        base = self.p_primary()
        if self.is_op("**"):
            self.p += 1
            # right associative; the exponent may carry a sign
            if self.is_op("-") or self.is_op("+"):
                op = self.peek().val; self.p += 1
                e = self.p_pow()
                if op == "-":
                    e = Node("un", op="neg", e=e)
            else:
                e = self.p_pow()
            return Node("bin", op="**", l=base, r=e)
        return base

    def p_primary(self):
        tk = self.peek()
        if tk is None:
            raise ParseError(f"unexpected end of expression in: {self.where}")
        if tk.kind == "int":
            self.p += 1
            return Node("num", k="i", v=tk.val)
        if tk.kind == "real":
            self.p += 1
            return Node("num", k="r8" if tk.kindp == "d" else "r4", v=tk.val)
        if tk.kind == "str":
            self.p += 1
            return Node("str", v=tk.val)
        if tk.kind == "log":
            self.p += 1
            return Node("log", v=tk.val)
        if tk.kind == "op" and tk.val == "(":
            self.p += 1
            e = self.expr()
            self.eat_op(")")
            return Node("paren", e=e)
        if tk.kind == "op" and tk.val in ("[", "(/"):
            close = "]" if tk.val == "[" else "/)"
            self.p += 1
            items = self.ac_items(close)
            self.eat_op(close)
            return Node("arr", items=items)
        if tk.kind == "id":
            return self.designator()
        raise ParseError(f"unexpected token {tk} in: {self.where}")

    def ac_items(self, close):
        items = []
        if self.is_op(close):
            return items
        while True:
            items.append(self.ac_item())
            if self.is_op(","):
                self.p += 1
                continue
            break
        return items

    def _implied_do_ahead(self):
        """At '(' : is this "( items , var = lo , hi [, st] )" ?"""
        depth, i = 0, self.p
        while i < len(self.t):
            tk = self.t[i]
            if tk.kind == "op" and tk.val in ("(", "[", "(/"):
                depth += 1
            elif tk.kind == "op" and tk.val in (")", "]", "/)"):
                depth -= 1
                if depth == 0:
                    return False
            elif depth == 1 and tk.kind == "op" and tk.val == "=" and self.t[i - 1].kind == "id" and self.t[i - 2].kind == "op" and self.t[i - 2].val == ",":
                return True
            i += 1
        return False

    def ac_item(self):
        if self.is_op("(") and self._implied_do_ahead():
            return self.implied_do()
        return self.expr()

    def implied_do(self):
        self.eat_op("(")
        items = []
        while True:
            if self.is_id() and self.is_op("=", 1):
                break
            items.append(self.ac_item())
            self.eat_op(",")
        var = self.eat_id()
        self.eat_op("=")
        lo = self.expr(); self.eat_op(",")
        hi = self.expr()
        st = None
        if self.is_op(","):
            self.p += 1
            st = self.expr()
        self.eat_op(")")
        return Node("ido", items=items, var=var, lo=lo, hi=hi, st=st)

    def args(self):
        """After '(' : argument / subscript list up to the matching ')'."""
        out = []
        if self.is_op(")"):
            self.p += 1
            return out
        while True:
            if self.is_id() and self.is_op("=", 1):
                name = self.eat_id(); self.p += 1
                out.append(Node("kw", name=name, e=self.expr()))
            else:
                lo = None
                if not self.is_op(":"):
                    lo = self.expr()
                if self.is_op(":"):
                    self.p += 1
                    hi = st = None
                    if not (self.is_op(",") or self.is_op(")") or self.is_op(":")):
                        hi = self.expr()
                    if self.is_op(":"):
                        self.p += 1
                        st = self.expr()
                    out.append(Node("slice", lo=lo, hi=hi, st=st))
                else:
                    out.append(lo)
            if self.is_op(","):
                self.p += 1
                continue
            self.eat_op(")")
            return out

    def designator(self):
        parts = []
        while True:
            name = self.eat_id()
            arglists = []
            while self.is_op("("):
                self.p += 1
                arglists.append(self.args())
            parts.append(Node("part", name=name, args=arglists[0] if arglists else None, sub=arglists[1] if len(arglists) > 1 else None))
            if self.is_op("%"):
                self.p += 1
                continue
            break
        return Node("desig", parts=parts)


# ---------------------------------------------------------------------------------------------------------
def split_top(toks: List[Tok], sep: str) -> List[List[Tok]]:
    out, cur, depth = [], [], 0
    for tk in toks:
        if tk.kind == "op" and tk.val in ("(", "[", "(/"):
            depth += 1
        elif tk.kind == "op" and tk.val in (")", "]", "/)"):
            depth -= 1
        if depth == 0 and tk.kind == "op" and tk.val == sep:
            out.append(cur); cur = []
        else:
            cur.append(tk)
    out.append(cur)
    return out


def match_paren(toks: List[Tok], i: int) -> int:
    """toks[i] is '(' : index of the matching ')'."""
    depth = 0
    for j in range(i, len(toks)):
        tk = toks[j]
        if tk.kind == "op" and tk.val in ("(", "(/"):
            depth += 1
        elif tk.kind == "op" and tk.val in (")", "/)"):
            depth -= 1
            if depth == 0:
                return j
    raise ParseError("unbalanced parentheses")


def parse_expr_toks(toks, where=""):
    ep = ExprParser(toks, 0, where)
    e = ep.expr()
    if not ep.at_end():
        raise ParseError(f"trailing tokens after expression in: {where}: {toks[ep.p:]}")
    return e


def is_assignment(toks: List[Tok]) -> bool:
    """designator '=' ... at depth 0, where the designator starts the statement."""
    i, n = 0, len(toks)
    if n < 3 or toks[0].kind != "id":
        return False
    while i < n:
        if toks[i].kind != "id":
            return False
        i += 1
        while i < n and toks[i].kind == "op" and toks[i].val == "(":
            i = match_paren(toks, i) + 1
        if i < n and toks[i].kind == "op" and toks[i].val == "%":
            i += 1
            continue
        break
    return i < n and toks[i].kind == "op" and toks[i].val in ("=", "=>")


# ---------------------------------------------------------------------------------------------------------
class Parser:
    def __init__(self, text: str, fname: str = ""):
        self.fname = fname
        self.lines = [(no, s, tokenize(s)) for no, s in logical_lines(text)]
        self.i = 0

    def cur(self):
        return self.lines[self.i] if self.i < len(self.lines) else None

    def err(self, msg):
        no, s, _ = self.cur() or (0, "<eof>", None)
        raise ParseError(f"{self.fname}:{no}: {msg}: {s}")

    # -- keyword helpers on a token list
    @staticmethod
    def kw(toks, *words):
        """Do the first len(words) tokens spell these identifiers?"""
        if len(toks) < len(words):
            return False
        return all(toks[k].kind == "id" and toks[k].val == w for k, w in enumerate(words))

    def is_end(self, toks, what):
        """'end what [name]' or 'endwhat [name]' or bare 'end' (for program units)."""
        if self.kw(toks, "end" + what):
            return True
        if self.kw(toks, "end", what):
            return True
        return False

    # -- file level
    def parse_file(self):
        units = []
        while self.cur():
            no, s, toks = self.cur()
            if self.kw(toks, "module") and not self.kw(toks, "module", "procedure"):
                units.append(self.parse_module())
            elif self.kw(toks, "program"):
                units.append(self.parse_proc("program"))
            elif self.proc_header(toks):
                units.append(self.parse_proc(self.proc_header(toks)))
            else:
                self.err("unexpected statement at file level")
        return units

    def proc_header(self, toks):
        """'subroutine' / 'function' if this statement opens one (with optional prefixes), else None."""
        k = 0
        while k < len(toks) and toks[k].kind == "id" and toks[k].val in ("recursive", "pure", "elemental"):
            k += 1
        if k < len(toks) and toks[k].kind == "id" and toks[k].val == "subroutine" and k + 1 < len(toks) and toks[k + 1].kind == "id":
            return "subroutine"
        # [type-spec] function name(
        j = k
        if j < len(toks) and toks[j].kind == "id" and toks[j].val in TYPE_WORDS:
            j += 1
            if j < len(toks) and toks[j].kind == "id" and toks[j].val == "precision":
                j += 1
            if j < len(toks) and toks[j].kind == "op" and toks[j].val == "(":
                j = match_paren(toks, j) + 1
            elif j < len(toks) and toks[j].kind == "op" and toks[j].val == "*":
                j += 2
        if j < len(toks) and toks[j].kind == "id" and toks[j].val == "function" and j + 2 < len(toks) and toks[j + 1].kind == "id" \
                and toks[j + 2].kind == "op" and toks[j + 2].val == "(":
            return "function"
        return None

    def parse_module(self):
        no, s, toks = self.cur()
        name = toks[1].val
        self.i += 1
        mod = Node("module", name=name, uses=[], decls=[], types=[], procs=[], interfaces=[], line=no)
        self.parse_spec(mod)
        no, s, toks = self.cur()
        if self.kw(toks, "contains"):
            self.i += 1
            while True:
                no, s, toks = self.cur()
                if self.is_end(toks, "module") or (len(toks) == 1 and self.kw(toks, "end")):
                    break
                h = self.proc_header(toks)
                if not h:
                    self.err("expected a procedure inside module")
                mod.procs.append(self.parse_proc(h))
        no, s, toks = self.cur()
        if not (self.is_end(toks, "module") or (len(toks) == 1 and self.kw(toks, "end"))):
            self.err("expected end module")
        self.i += 1
        return mod

    def parse_spec(self, unit):
        """use / implicit / access / declarations / type definitions, until the first other statement."""
        while self.cur():
            no, s, toks = self.cur()
            if self.kw(toks, "use"):
                # use name | use name, only: ... | use, intrinsic :: name
                dc = next((j for j, tk in enumerate(toks) if tk.kind == "op" and tk.val == "::"), None)
                unit.uses.append(toks[dc + 1].val if dc is not None else toks[1].val)
                self.i += 1
            elif self.kw(toks, "import"):
                self.i += 1
            elif self.kw(toks, "interface") and len(toks) == 1:
                unit.interfaces.extend(self.parse_interface())
            elif self.kw(toks, "implicit"):
                self.i += 1
            elif (self.kw(toks, "private") or self.kw(toks, "public")) and not is_assignment(toks):
                self.i += 1
            elif self.kw(toks, "type") and not (len(toks) > 1 and toks[1].kind == "op" and toks[1].val == "(") and not is_assignment(toks):
                unit.types.append(self.parse_typedef())
            elif toks[0].kind == "id" and toks[0].val in TYPE_WORDS and not is_assignment(toks) and not self.proc_header(toks):
                unit.decls.append(self.parse_decl(toks, no, s))
                self.i += 1
            elif self.kw(toks, "save") and len(toks) == 1:
                self.i += 1
            else:
                return

    def parse_typedef(self):
        no, s, toks = self.cur()
        # type [, attrs] [::] name
        name = toks[-1].val
        self.i += 1
        td = Node("typedef", name=name, comps=[], bindings={}, line=no)
        in_contains = False
        while True:
            no, s, toks = self.cur()
            if self.is_end(toks, "type"):
                self.i += 1
                return td
            if self.kw(toks, "contains"):
                in_contains = True
                self.i += 1
                continue
            if in_contains:
                # procedure [, attrs] :: binding => target [, ...]
                if not self.kw(toks, "procedure"):
                    self.err("expected type-bound procedure")
                k = next(j for j, tk in enumerate(toks) if tk.kind == "op" and tk.val == "::")
                for grp in split_top(toks[k + 1:], ","):
                    b = grp[0].val
                    tgt = grp[2].val if len(grp) >= 3 else b
                    td.bindings[b] = tgt
                self.i += 1
                continue
            if self.kw(toks, "private") or self.kw(toks, "public") or self.kw(toks, "sequence"):
                self.i += 1
                continue
            td.comps.append(self.parse_decl(toks, no, s))
            self.i += 1

    def parse_type_spec(self, toks, k, where):
        """toks[k:] starts with a type-spec; returns (spec Node, next index)."""
        w = toks[k].val
        k += 1
        spec = Node("tspec", base=w, kind=None, len=None, tname=None)
        if w == "double":
            k += 1   # precision
            spec.base, spec.kind = "real", Node("num", k="i", v="8")
            return spec, k
        if w in ("type", "class"):
            j = match_paren(toks, k)
            spec.base, spec.tname = "type", toks[k + 1].val
            return spec, j + 1
        if k < len(toks) and toks[k].kind == "op" and toks[k].val == "*":      # real*8, character*40
            v = Node("num", k="i", v=toks[k + 1].val)
            if w == "character":
                spec.len = v
            else:
                spec.kind = v
            return spec, k + 2
        if k < len(toks) and toks[k].kind == "op" and toks[k].val == "(":
            j = match_paren(toks, k)
            inner = toks[k + 1:j]
            for grp in split_top(inner, ","):
                key = None
                if len(grp) >= 2 and grp[0].kind == "id" and grp[1].kind == "op" and grp[1].val == "=":
                    key, grp = grp[0].val, grp[2:]
                if len(grp) == 1 and grp[0].kind == "op" and grp[0].val in ("*", ":"):
                    val = Node("star")
                else:
                    val = parse_expr_toks(grp, where)
                if w == "character":
                    if key in (None, "len"):
                        spec.len = val
                else:
                    spec.kind = val
            return spec, j + 1
        return spec, k

    def parse_decl(self, toks, no, s):
        spec, k = self.parse_type_spec(toks, 0, s)
        attrs = {}
        # attributes up to '::' (if there is one)
        dc = next((j for j, tk in enumerate(toks) if tk.kind == "op" and tk.val == "::"), None)
        if dc is not None:
            for grp in split_top(toks[k:dc], ","):
                if not grp:
                    continue
                a = grp[0].val
                if a == "dimension":
                    attrs["dimension"] = self.parse_dims(grp[2:-1], s)
                elif a == "intent":
                    attrs["intent"] = "".join(t.val for t in grp[2:-1])
                else:
                    attrs[a] = True
            k = dc + 1
        ents = []
        for grp in split_top(toks[k:], ","):
            if not grp:
                continue
            name = grp[0].val
            j = 1
            dims = None
            clen = None
            init = None
            if j < len(grp) and grp[j].kind == "op" and grp[j].val == "(":
                e = match_paren(grp, j)
                dims = self.parse_dims(grp[j + 1:e], s)
                j = e + 1
            if j < len(grp) and grp[j].kind == "op" and grp[j].val == "*":
                clen = Node("num", k="i", v=grp[j + 1].val)
                j += 2
            if j < len(grp) and grp[j].kind == "op" and grp[j].val == "=":
                init = parse_expr_toks(grp[j + 1:], s)
            ents.append(Node("entity", name=name, dims=dims, clen=clen, init=init))
        return Node("decl", spec=spec, attrs=attrs, ents=ents, line=no)

    def parse_dims(self, toks, where):
        dims = []
        for grp in split_top(toks, ","):
            parts = split_top(grp, ":")
            if len(parts) == 1:
                if len(parts[0]) == 1 and parts[0][0].kind == "op" and parts[0][0].val == "*":
                    dims.append((Node("num", k="i", v="1"), Node("star")))
                else:
                    dims.append((Node("num", k="i", v="1"), parse_expr_toks(parts[0], where)))
            else:
                lo = parse_expr_toks(parts[0], where) if parts[0] else None
                hi = None
                if parts[1]:
                    if len(parts[1]) == 1 and parts[1][0].kind == "op" and parts[1][0].val == "*":
                        hi = Node("star")
                    else:
                        hi = parse_expr_toks(parts[1], where)
                dims.append((lo, hi))   # (None, None) = deferred / assumed shape
        return dims

    # -- procedures
    def parse_proc(self, kind):
        no, s, toks = self.cur()
        k = 0
        while toks[k].val in ("recursive", "pure", "elemental"):
            k += 1
        rtype = None
        if kind == "function" and toks[k].val != "function":
            rtype, k = self.parse_type_spec(toks, k, s)
        k += 1   # subroutine / function / program
        name = toks[k].val
        k += 1
        args = []
        result = None
        if k < len(toks) and toks[k].kind == "op" and toks[k].val == "(":
            e = match_paren(toks, k)
            args = [g[0].val for g in split_top(toks[k + 1:e], ",") if g]
            k = e + 1
        if k < len(toks) and toks[k].kind == "id" and toks[k].val == "result":
            result = toks[k + 2].val
        self.i += 1
        cname = None
        bi = next((j for j in range(k, len(toks)) if toks[j].kind == "id" and toks[j].val == "bind"), None)
        if bi is not None:      # bind(C [, name='...'])
            e = match_paren(toks, bi + 1)
            cname = name
            for j in range(bi + 2, e):
                if toks[j].kind == "str":
                    cname = toks[j].val
        proc = Node("proc", kind=kind, name=name, args=args, result=result, rtype=rtype, uses=[], decls=[], types=[], body=[], procs=[], line=no,
                    file=self.fname, interfaces=[], cname=cname)
        self.parse_spec(proc)
        proc.body = self.parse_block(("contains", "end"))
        no, s, toks = self.cur()
        if self.kw(toks, "contains"):
            self.i += 1
            while True:
                no, s, toks = self.cur()
                h = self.proc_header(toks)
                if not h:
                    break
                proc.procs.append(self.parse_proc(h))
        no, s, toks = self.cur()
        if not (self.kw(toks, "end") or self.kw(toks, "end" + kind)):
            self.err(f"expected end of {kind} {name}")
        self.i += 1
        return proc

    def parse_interface(self):
        """interface ... end interface: explicit interfaces of external (here: bind(C)) procedures."""
        self.i += 1
        out = []
        while True:
            no, s, toks = self.cur()
            if self.kw(toks, "end", "interface") or self.kw(toks, "endinterface"):
                self.i += 1
                return out
            h = self.proc_header(toks)
            if not h:
                self.err("expected a procedure interface")
            out.append(self.parse_proc(h))

    def block_end(self, toks, enders):
        """Does this statement close the current block?  enders: tuple of words; 'end' matches 'end xxx' / 'endxxx'."""
        w = toks[0].val if toks[0].kind == "id" else None
        if w is None:
            return False
        if is_assignment(toks):
            return False
        for e in enders:
            if e == "end":
                if w == "end" or (w.startswith("end") and w[3:] in ("subroutine", "function", "program", "module")):
                    if w == "end" and len(toks) > 1 and toks[1].val in ("if", "do", "select", "associate", "where", "type"):
                        continue
                    return True
            elif e == "contains":
                if w == "contains" and len(toks) == 1:
                    return True
            elif e == "enddo":
                if w == "enddo" or (w == "end" and len(toks) > 1 and toks[1].val == "do"):
                    return True
            elif e == "endif":
                if w == "endif" or (w == "end" and len(toks) > 1 and toks[1].val == "if"):
                    return True
            elif e == "else":
                if w in ("else", "elseif"):
                    return True
            elif e == "case":
                if w == "case":
                    return True
            elif e == "endselect":
                if w == "endselect" or (w == "end" and len(toks) > 1 and toks[1].val == "select"):
                    return True
            elif e == "endassociate":
                if w == "endassociate" or (w == "end" and len(toks) > 1 and toks[1].val == "associate"):
                    return True
        return False

    def parse_block(self, enders):
        stmts = []
        while self.cur():
            no, s, toks = self.cur()
            if self.block_end(toks, enders):
                return stmts
            stmts.append(self.parse_stmt())
        self.err("unexpected end of file in block")

    def parse_stmt(self):
        no, s, toks = self.cur()
        st = self.parse_simple(toks, no, s, allow_block=True)
        return st

    def parse_simple(self, toks, no, s, allow_block=False):
        """One statement; block constructs consume further lines (only when allow_block)."""
        if is_assignment(toks):
            k = next(j for j in range(len(toks)) if toks[j].kind == "op" and toks[j].val in ("=", "=>") and self._depth0(toks, j))
            lhs = parse_expr_toks(toks[:k], s)
            rhs = parse_expr_toks(toks[k + 1:], s)
            if allow_block:
                self.i += 1
            return Node("assign", lhs=lhs, rhs=rhs, line=no)
        w = toks[0].val if toks[0].kind == "id" else None
        if w == "if" or w == "elseif":
            e = match_paren(toks, 1)
            cond = parse_expr_toks(toks[2:e], s)
            rest = toks[e + 1:]
            if len(rest) == 1 and rest[0].kind == "id" and rest[0].val == "then":
                if not allow_block:
                    self.err("block if not allowed here")
                return self.parse_if_block(cond, no)
            inner = self.parse_simple(rest, no, s, allow_block=False)
            if allow_block:
                self.i += 1
            return Node("if", branches=[(cond, [inner])], orelse=None, line=no)
        if w == "do":
            if not allow_block:
                self.err("do not allowed here")
            return self.parse_do(toks, no, s)
        if w == "select":
            return self.parse_select(toks, no, s)
        if w == "associate":
            return self.parse_associate(toks, no, s)
        st = None
        if w == "call":
            ep = ExprParser(toks, 1, s)
            d = ep.designator()
            st = Node("call", target=d, line=no)
        elif w in ("cycle", "exit", "return", "continue"):
            st = Node(w, line=no)
        elif w == "stop":
            st = Node("stop", msg=parse_expr_toks(toks[1:], s) if len(toks) > 1 else None, line=no)
        elif w in ("allocate", "deallocate"):
            e = match_paren(toks, 1)
            items, stat = [], None
            ep = ExprParser(toks[2:e], 0, s)
            for a in ep.args_noparen():
                if a.t == "kw":
                    stat = a
                else:
                    items.append(a)
            st = Node(w, items=items, line=no)
        elif w in ("open", "close", "read", "write", "rewind", "flush", "inquire", "backspace"):
            st = self.parse_io(w, toks, no, s)
        elif w == "print":
            fmt = toks[1]
            items = self.parse_io_items(toks[3:], s) if len(toks) > 2 else []
            ctl = [Node("star")] if (fmt.kind == "op" and fmt.val == "*") else [parse_expr_toks([fmt], s)]
            st = Node("write", ctl=[Node("star")] + ctl, kws={}, items=items, line=no)
        if st is None:
            self.err("cannot parse statement")
        if allow_block:
            self.i += 1
        return st

    @staticmethod
    def _depth0(toks, j):
        d = 0
        for tk in toks[:j]:
            if tk.kind == "op" and tk.val in ("(", "[", "(/"):
                d += 1
            elif tk.kind == "op" and tk.val in (")", "]", "/)"):
                d -= 1
        return d == 0

    def parse_if_block(self, cond, no):
        self.i += 1
        branches = []
        orelse = None
        body = self.parse_block(("else", "endif"))
        branches.append((cond, body))
        while True:
            ln, s, toks = self.cur()
            w = toks[0].val
            if w == "elseif" or (w == "else" and len(toks) > 1 and toks[1].kind == "id" and toks[1].val == "if"):
                k = 1 if w == "elseif" else 2
                e = match_paren(toks, k)
                c = parse_expr_toks(toks[k + 1:e], s)
                self.i += 1
                body = self.parse_block(("else", "endif"))
                branches.append((c, body))
            elif w == "else":
                self.i += 1
                orelse = self.parse_block(("endif",))
            else:   # endif
                self.i += 1
                return Node("if", branches=branches, orelse=orelse, line=no)

    def parse_do(self, toks, no, s):
        self.i += 1
        if len(toks) == 1:
            body = self.parse_block(("enddo",))
            self.i += 1
            return Node("dowhile", cond=Node("log", v=True), body=body, line=no)
        if toks[1].kind == "id" and toks[1].val == "while":
            e = match_paren(toks, 2)
            cond = parse_expr_toks(toks[3:e], s)
            body = self.parse_block(("enddo",))
            self.i += 1
            return Node("dowhile", cond=cond, body=body, line=no)
        var = toks[1].val
        parts = split_top(toks[3:], ",")
        lo = parse_expr_toks(parts[0], s); hi = parse_expr_toks(parts[1], s)
        st = parse_expr_toks(parts[2], s) if len(parts) > 2 else None
        body = self.parse_block(("enddo",))
        self.i += 1
        return Node("do", var=var, lo=lo, hi=hi, st=st, body=body, line=no)

    def parse_select(self, toks, no, s):
        # select case (expr)
        e = match_paren(toks, 2)
        sel = parse_expr_toks(toks[3:e], s)
        self.i += 1
        cases = []
        default = None
        while True:
            ln, s2, tk = self.cur()
            if self.block_end(tk, ("endselect",)):
                self.i += 1
                return Node("select", sel=sel, cases=cases, default=default, line=no)
            if tk[0].val != "case":
                self.err("expected case")
            self.i += 1
            if len(tk) > 1 and tk[1].kind == "id" and tk[1].val == "default":
                default = self.parse_block(("case", "endselect"))
            else:
                ce = match_paren(tk, 1)
                vals = []
                for grp in split_top(tk[2:ce], ","):
                    rng = split_top(grp, ":")
                    if len(rng) == 1:
                        vals.append(("v", parse_expr_toks(rng[0], s2)))
                    else:
                        vals.append(("r", parse_expr_toks(rng[0], s2) if rng[0] else None, parse_expr_toks(rng[1], s2) if rng[1] else None))
                body = self.parse_block(("case", "endselect"))
                cases.append((vals, body))

    def parse_associate(self, toks, no, s):
        e = match_paren(toks, 1)
        pairs = []
        for grp in split_top(toks[2:e], ","):
            pairs.append((grp[0].val, parse_expr_toks(grp[2:], s)))
        self.i += 1
        body = self.parse_block(("endassociate",))
        self.i += 1
        return Node("associate", pairs=pairs, body=body, line=no)

    def parse_io(self, w, toks, no, s):
        if toks[1].kind == "op" and toks[1].val == "(":
            e = match_paren(toks, 1)
            ctl, kws = [], {}
            for grp in split_top(toks[2:e], ","):
                if len(grp) >= 2 and grp[0].kind == "id" and grp[1].kind == "op" and grp[1].val == "=":
                    kws[grp[0].val] = Node("star") if (len(grp) == 3 and grp[2].kind == "op" and grp[2].val == "*") else parse_expr_toks(grp[2:], s)
                elif len(grp) == 1 and grp[0].kind == "op" and grp[0].val == "*":
                    ctl.append(Node("star"))
                else:
                    ctl.append(parse_expr_toks(grp, s))
            items = self.parse_io_items(toks[e + 1:], s) if e + 1 < len(toks) else []
            return Node(w, ctl=ctl, kws=kws, items=items, line=no)
        # "rewind 111", "read *, x" forms
        return Node(w, ctl=[parse_expr_toks(toks[1:2], s)], kws={}, items=[], line=no)

    def parse_io_items(self, toks, s):
        items = []
        for grp in split_top(toks, ","):
            if not grp:
                continue
            ep = ExprParser(grp, 0, s)
            items.append(ep.ac_item())
            if not ep.at_end():
                raise ParseError(f"bad I/O item in: {s}")
        return items


def _args_noparen(self):
    """Argument list without the surrounding parentheses (allocate lists)."""
    out = []
    while not self.at_end():
        if self.is_id() and self.is_op("=", 1):
            name = self.eat_id(); self.p += 1
            out.append(Node("kw", name=name, e=self.expr()))
        else:
            out.append(self.designator())
        if self.is_op(","):
            self.p += 1
    return out


ExprParser.args_noparen = _args_noparen


def parse_source(text: str, fname: str = ""):
    return Parser(text, fname).parse_file()

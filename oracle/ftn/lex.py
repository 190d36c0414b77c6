"""Free-form Fortran source -> logical statements -> tokens.

TEST INFRASTRUCTURE (part of oracle/): the front end of the small Fortran interpreter that executes the reference's own
source files (read from /root/reference at run time, never copied) so that the C oracle can be pinned against the reference
itself in an image that has no Fortran compiler.  See oracle/ftn/README.md.
"""
from __future__ import annotations

import re
from typing import List, NamedTuple, Tuple


class Tok(NamedTuple):
    kind: str     # 'id', 'int', 'real', 'str', 'op', 'dotop', 'log'
    val: object   # identifier (lower case), number text, string contents, operator text
    kindp: str = ""   # for 'real': 'd' (double: d exponent or _8 suffix) or 's' (default real)


DOT_WORDS = ("eq", "ne", "lt", "le", "gt", "ge", "and", "or", "not", "eqv", "neqv", "true", "false")
_dot_re = re.compile(r"\.(" + "|".join(sorted(DOT_WORDS, key=len, reverse=True)) + r")\.", re.I)
_id_re = re.compile(r"[A-Za-z_][A-Za-z0-9_]*")
_num_re = re.compile(r"(\d+\.?\d*|\.\d+)([eEdD][+-]?\d+)?(_\w+)?")
OPS3 = ()
OPS2 = ("**", "//", "==", "/=", "<=", ">=", "=>", "::", "(/", "/)")
OPS1 = "+-*/=<>(),:%[]"


def strip_comment(line: str) -> str:
    """Removes a trailing ! comment (not inside a character literal)."""
    q = None
    for i, c in enumerate(line):
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c == "!":
            return line[:i]
    return line


def logical_lines(text: str) -> List[Tuple[int, str]]:
    """Joins continuation lines, drops comments and blank lines, splits on ';'.  Returns (first physical line number, text)."""
    out: List[Tuple[int, str]] = []
    cur, cur_no = "", 0
    for no, raw in enumerate(text.splitlines(), 1):
        line = strip_comment(raw).rstrip()
        s = line.strip()
        if not s:
            continue
        if cur:
            if s.startswith("&"):
                s = s[1:]
            piece = s
        else:
            piece = s
            cur_no = no
        if piece.endswith("&"):
            cur += piece[:-1]
            continue
        cur += piece
        # split on ';' outside strings
        q, start = None, 0
        for i, c in enumerate(cur):
            if q:
                if c == q:
                    q = None
            elif c in "'\"":
                q = c
            elif c == ";":
                if cur[start:i].strip():
                    out.append((cur_no, cur[start:i].strip()))
                start = i + 1
        if cur[start:].strip():
            out.append((cur_no, cur[start:].strip()))
        cur = ""
    if cur.strip():
        out.append((cur_no, cur.strip()))
    return out


def tokenize(s: str) -> List[Tok]:
    toks: List[Tok] = []
    i, n = 0, len(s)
    while i < n:
        c = s[i]
        if c in " \t":
            i += 1
            continue
        if c in "'\"":
            j, buf = i + 1, []
            while j < n:
                if s[j] == c:
                    if j + 1 < n and s[j + 1] == c:
                        buf.append(c); j += 2
                        continue
                    break
                buf.append(s[j]); j += 1
            if j >= n:
                raise SyntaxError(f"unterminated string in: {s}")
            toks.append(Tok("str", "".join(buf)))
            i = j + 1
            continue
        if c == ".":
            m = _dot_re.match(s, i)
            if m:
                w = m.group(1).lower()
                if w in ("true", "false"):
                    toks.append(Tok("log", w == "true"))
                else:
                    toks.append(Tok("dotop", w))
                i = m.end()
                continue
        if c.isdigit() or (c == "." and i + 1 < n and s[i + 1].isdigit()):
            m = _num_re.match(s, i)
            txt = m.group(0)
            mant, exp, suf = m.group(1), m.group(2), m.group(3)
            # "1.eq.2": the dot belongs to the operator, not to the number
            if mant.endswith(".") and not exp and _dot_re.match(s, i + len(mant) - 1):
                mant = mant[:-1]
                txt = mant
                exp = suf = None
            if "." in mant or exp:
                dbl = bool(exp and exp[0] in "dD") or (suf is not None and suf[1:] in ("8", "dp"))
                val = mant + (("e" + exp[1:]) if exp else "")
                toks.append(Tok("real", val, "d" if dbl else "s"))
            else:
                toks.append(Tok("int", mant, (suf or "")[1:]))
            i += len(txt)
            continue
        m = _id_re.match(s, i)
        if m:
            toks.append(Tok("id", m.group(0).lower()))
            i = m.end()
            continue
        two = s[i:i + 2]
        if two in OPS2:
            # "(/" opens an array constructor only when not followed by "=" (then it is "(" and "/=")
            if two == "(/" and s[i + 2:i + 3] == "=":
                toks.append(Tok("op", "(")); i += 1
                continue
            toks.append(Tok("op", two)); i += 2
            continue
        if c in OPS1:
            toks.append(Tok("op", c)); i += 1
            continue
        raise SyntaxError(f"unexpected character {c!r} in: {s}")
    return toks

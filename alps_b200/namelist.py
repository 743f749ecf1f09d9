"""Reader for the reference's Fortran-namelist inputs (`*.in`, `*_dist.in`): groups `&name ... /`
with `key=value` entries, `!` comments, Fortran literals (1.d-2, .true., T, 'text').
Indexed groups (&spec_1, &ffit_1_1, &scan_input_1, ...) are returned under their full name, as
get_indexed_namelist_unit does (src/ALPS_io.f90:698-823)."""
from __future__ import annotations

import re
from typing import Dict


def _value(tok: str):
    t = tok.strip().rstrip(",").strip()
    if not t:
        return None
    if t[0] in "'\"":
        return t.strip("'\"")
    low = t.lower()
    if low in (".true.", "t", ".t."):
        return True
    if low in (".false.", "f", ".f."):
        return False
    try:
        return int(t)
    except ValueError:
        pass
    try:
        return float(low.replace("d", "e"))
    except ValueError:
        return t


def read_namelists(path: str) -> Dict[str, Dict[str, object]]:
    groups: Dict[str, Dict[str, object]] = {}
    cur = None
    for raw in open(path):
        line = raw
        # strip comments (a '!' outside quotes)
        out, q = [], None
        for ch in line:
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "!":
                break
            out.append(ch)
        line = "".join(out).strip()
        if not line:
            continue
        if line.startswith("&"):
            cur = line[1:].split()[0].lower()
            groups[cur] = {}
            line = line[1 + len(cur):].strip()
            if not line:
                continue
        if line == "/" or line.lower() == "&end":
            cur = None
            continue
        if cur is None:
            continue
        if line.endswith("/"):
            body, end = line[:-1], True
        else:
            body, end = line, False
        for m in re.finditer(r"(\w+)\s*=\s*('[^']*'|\"[^\"]*\"|[^,\s]+)", body):
            groups[cur][m.group(1).lower()] = _value(m.group(2))
        if end:
            cur = None
    return groups

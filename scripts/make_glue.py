#!/usr/bin/env python
"""Cut the four step routines out of the model's advance.f, leaving the glue that stays Fortran.

    python scripts/make_glue.py pom/advance.f pom/advance_glue.f

libpomgpu_f.so exports `lateral_viscosity_`, `mode_interaction_`, `mode_external_`, `mode_internal_`
(pom/advance.f:96,144,205,356).  A definition inside the executable would take precedence over the
shared library's, so the GPU build must not compile the Fortran bodies of those four; everything
else in advance.f (`advance`, `get_time`, `surface_forcing`, `print_section`, `check_velocity`,
`domain_stats`) is driver glue and stays.  This script copies advance.f without the four bodies --
no line of the remaining routines is edited -- and the makefile lists advance_glue.o instead of
advance.o and drops solver.o (every routine of solver.f is provided by the library).  See
INTEGRATION.md."""
import re
import sys

CUT = ("lateral_viscosity", "mode_interaction", "mode_external", "mode_internal")


def cut(src):
    out, skipping, removed = [], None, []
    for line in src.splitlines(keepends=True):
        code = line.split("!")[0] if not line[:1] in "cC*" else ""
        m = re.match(r"\s+subroutine\s+(\w+)", code, re.I)
        if skipping is None and m and m.group(1).lower() in CUT:
            skipping = m.group(1).lower()
            removed.append(skipping)
            out.append(f"! [{skipping}: provided by libpomgpu_f.so]\n")
            continue
        if skipping is not None:
            if re.match(r"\s+end(\s+subroutine(\s+\w+)?)?\s*$", code, re.I):
                skipping = None
            continue
        out.append(line)
    return "".join(out), removed


if __name__ == "__main__":
    text, removed = cut(open(sys.argv[1]).read())
    assert sorted(removed) == sorted(CUT), f"found {removed}"
    open(sys.argv[2], "w").write(text)
    print(f"{sys.argv[2]}: removed {', '.join(removed)}")

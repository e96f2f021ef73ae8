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
INTEGRATION.md.

    python scripts/make_glue.py pom/advance.f pom/advance_glue.f pom/bounds_forcing.f

With the third argument the glue file also gets `subroutine restore_interior_records`: the record half of
`restore_interior` (pom/bounds_forcing.f:1023-1081 -- declarations, the netCDF reads, the `b = f` copies), i.e. the
routine's own lines up to the comment "linear interpolation in time", closed with `return` / `end`; only the name on
the `subroutine` line differs.  The reference calls restore_interior from INSIDE mode_internal (advance.f:452), which is
the library's now: libpomgpu_f's mode_internal_ calls this routine back at that place and does the interpolation and
the nudging (:1083-1118) on the device."""
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


def restore_records(src):
    """`subroutine restore_interior_records` from the text of bounds_forcing.f (see the module docstring)."""
    out, inside = [], False
    for line in src.splitlines(keepends=True):
        code = line.split("!")[0] if not line[:1] in "cC*" else ""
        if not inside:
            m = re.match(r"(\s+subroutine\s+)restore_interior\b(.*)$", code, re.I)
            if m:
                inside = True
                out.append("! [the record half of restore_interior (bounds_forcing.f), called back by libpomgpu_f.so's mode_internal_]\n")
                out.append(f"{m.group(1)}restore_interior_records\n")
            continue
        if re.match(r"\s*!\s*linear interpolation in time", line, re.I):
            out.append("      return\n      end\n")
            return "".join(out)
        if re.match(r"\s+end(\s+subroutine(\s+\w+)?)?\s*$", code, re.I):
            break
        out.append(line)
    raise ValueError("restore_interior / its 'linear interpolation in time' comment not found")


if __name__ == "__main__":
    text, removed = cut(open(sys.argv[1]).read())
    assert sorted(removed) == sorted(CUT), f"found {removed}"
    if len(sys.argv) > 3:
        text += "\n!_______________________________________________________________________\n" + restore_records(open(sys.argv[3]).read())
    open(sys.argv[2], "w").write(text)
    print(f"{sys.argv[2]}: removed {', '.join(removed)}" + ("; added restore_interior_records" if len(sys.argv) > 3 else ""))

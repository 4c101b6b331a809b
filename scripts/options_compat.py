"""Which of the solver option files PetIBM ships (examples/**/config/*_solver.info, *.info) this backend accepts as they
are, and why not otherwise.  Run where the reference checkout is available:

    python scripts/options_compat.py /root/reference/examples > /tmp/compat.md
"""
import ctypes as C
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petibm_b200 import _lib  # noqa: E402


def main(root):
    L = _lib.lib()
    rows = {}
    for dirpath, _, files in os.walk(root):
        for f in sorted(files):
            if not f.endswith(".info"):
                continue
            text = open(os.path.join(dirpath, f)).read()
            if "config_version" in text:                       # AmgX JSON/ini configuration, not PETSc options
                verdict = "AmgX configuration file (backend `GPU`): not PETSc options"
                key = (f, "amgx", verdict)
            else:
                m = re.search(r"^-([a-z]+)_", text, re.M)
                prefix = (m.group(1) if m else f.split("_")[0]) + "_"
                o = _lib.Options()
                L.b200ls_default_options(C.byref(o))
                err = C.create_string_buffer(256)
                rc = L.b200ls_parse_options(text.encode(), prefix.encode(), C.byref(o), err, 256)
                opts = " ".join(l.strip() for l in text.splitlines() if l.strip().startswith("-"))
                verdict = "accepted" if rc == 0 else "refused: " + err.value.decode()
                key = (f, opts, verdict)
            rows.setdefault(key, []).append(os.path.relpath(dirpath, root))
    print("| file | options | verdict | used by |")
    print("|---|---|---|---|")
    for (f, opts, verdict), where in sorted(rows.items()):
        cases = sorted({w.split(os.sep + "config")[0] for w in where})
        shown = ", ".join(cases[:3]) + (f", … ({len(cases)} cases)" if len(cases) > 3 else "")
        print(f"| `{f}` | `{opts if opts != 'amgx' else '(AmgX)'}` | {verdict} | {shown} |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/examples")

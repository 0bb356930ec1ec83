#!/usr/bin/env python3
"""Writes tests/golden/loudgain_deb_imports.json: what the reference's shipped binary
(/root/reference/bin/loudgain_0.5.3-1ubuntu1_amd64.deb -> usr/bin/loudgain) imports from
libebur128 -- its undefined ebur128_* symbols and the SONAME it NEEDs.  Run in the
build container (the reference tree does not travel to the GPU box); the fixture does."""
import json
import os
import subprocess
import sys
import tempfile

DEB = "/root/reference/bin/loudgain_0.5.3-1ubuntu1_amd64.deb"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def deb_imports(deb=DEB):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call(["ar", "x", deb], cwd=tmp)
        data = [f for f in os.listdir(tmp) if f.startswith("data.tar")][0]
        subprocess.check_call(["tar", "xf", data, "./usr/bin/loudgain"], cwd=tmp)
        exe = os.path.join(tmp, "usr/bin/loudgain")
        nm = subprocess.check_output(["nm", "-D", "--undefined-only", exe], text=True)
        syms = sorted(line.split()[-1] for line in nm.splitlines() if "ebur128_" in line)
        dyn = subprocess.check_output(["readelf", "-d", exe], text=True)
        needed = [line.split("[")[1].split("]")[0] for line in dyn.splitlines() if "(NEEDED)" in line]
    return {"source": os.path.basename(deb) + ": usr/bin/loudgain (nm -D --undefined-only, readelf -d)",
            "undefined_ebur128_symbols": syms,
            "needed": needed}


if __name__ == "__main__":
    out = deb_imports()
    path = os.path.join(ROOT, "tests", "golden", "loudgain_deb_imports.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))
    sys.exit(0)

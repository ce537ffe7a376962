"""Development aid: build variants of the library with extra -D switches and time erode(512) with each.
usage: python tools/variants.py MAPSIZE name:-DA,-DB ...   (name "base" with no defines = the shipped build)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplehydrology_b200 import build as B  # noqa: E402


def build_variant(name, defs):
    lib = os.path.join(ROOT, "simplehydrology_b200", "_variants", f"libshx_var_{name}.so")
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    r = subprocess.run([B.nvcc()] + B.NVCC_FLAGS + defs + ["-Xptxas", "-v", "-o", lib] + B.SOURCES, check=True, capture_output=True, text=True)
    lines = r.stderr.splitlines()
    for i, l in enumerate(lines):  # the dense launch shape
        if "descend_lockstep_kernelILi448ELi2ELb1ELb0" in l and "Function properties" in l:
            print("   ", lines[i + 1].strip(), "|", lines[i + 2].strip())
    return lib


if __name__ == "__main__":
    if sys.argv[1] == "build":
        for spec in sys.argv[2:]:
            name, _, d = spec.partition(":")
            print(build_variant(name, [x for x in d.split(",") if x]))
    else:
        ms = sys.argv[1]
        for spec in sys.argv[2:]:
            name = spec.partition(":")[0]
            lib = os.path.join(ROOT, "simplehydrology_b200", "_variants", f"libshx_var_{name}.so")
            env = dict(os.environ, SHX_LIB=lib)
            print(f"== variant {name}", flush=True)
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "tune_descend.py"), ms, "0:0:0"], env=env)

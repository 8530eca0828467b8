"""Development tool: build compile-time variants of libtvf.so into tools/_build/variants/ so that ONE GPU call can
time them all (TVF_LIBPATH selects the library bench.py loads).  Not part of the product build.

    python tools/build_variants.py name=-DFLAG=1,-DOTHER=2 name2=...
"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tft_vs_fund_b200 import build as B   # noqa: E402

OUT = os.path.join(ROOT, "tools", "_build", "variants")


def build_one(spec):
    name, _, flags = spec.partition("=")
    defines = [f for f in flags.split(",") if f]
    out = os.path.join(OUT, "libtvf_%s.so" % name)
    os.makedirs(OUT, exist_ok=True)
    try:
        B.build(force=True, defines=defines, out=out)
        return name, "ok"
    except Exception as e:
        return name, "FAILED: %s" % e


if __name__ == "__main__":
    specs = sys.argv[1:] or ["base="]
    with ThreadPoolExecutor(3) as ex:
        for name, msg in ex.map(build_one, specs):
            print(name, msg)

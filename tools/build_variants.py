"""Development tool: build compile-time variants of libtvf.so into tools/_build/variants/ so that ONE GPU call can
time them all (TVF_LIBPATH selects the library bench.py loads).  Not part of the product build."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tft_vs_fund_b200 import build as B   # noqa: E402

VARIANTS = {
    "base": [],
    "nopipe": ["-DTVF_GJ_PIPE=0"],
}


def build_one(name):
    out = os.path.join(ROOT, "tools", "_build", "variants", "libtvf_%s.so" % name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [B._nvcc()] + [f for f in B.NVCC_FLAGS if f != "-Xptxas=-v"] + VARIANTS[name] + ["-I", B.CSRC, "-o", out] + \
        [os.path.join(B.CSRC, f) for f in B.SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return name, res.returncode, (res.stdout + res.stderr)[-2000:]


if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    with ThreadPoolExecutor(4) as ex:
        for name, rc, log in ex.map(build_one, names):
            print(name, "ok" if rc == 0 else "FAILED\n" + log)

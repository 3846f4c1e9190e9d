#!/usr/bin/env python
"""Compile the reference's own Cython segmentation kernels
(/root/reference/warpdemux/segmentation/_c_segmentation.pyx) from where they
lie into oracle/_ref/ (git-ignored, NOT gpurun-ignored).  Test infrastructure:
used to validate the restated windowed t-test / segment means against the
reference's native code.  No reference source is copied into the repo — only
the generated C++ and the built extension land in oracle/_ref/."""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WDX_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "warpdemux", "segmentation", "_c_segmentation.pyx")
OUT = os.path.join(HERE, "_ref")


def _build_one(src: str, modname: str) -> None:
    import numpy as np

    ext = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(OUT, modname + ext)
    if os.path.exists(target) and os.path.getmtime(target) >= os.path.getmtime(src):
        print("up to date:", target)
        return
    cpp = os.path.join(OUT, modname + ".cpp")
    # cython needs the module name to match the file: generate under the target name
    subprocess.check_call([sys.executable, "-m", "cython", "--cplus", "-3", "--module-name", modname, src, "-o", cpp])
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
    cmd = [cxx, "-O2", "-fPIC", "-shared", "-std=c++17", "-fno-fast-math", "-ffp-contract=off",
           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
           "-I", sysconfig.get_paths()["include"], "-I", np.get_include(), cpp, "-o", target]
    subprocess.check_call(cmd)
    os.remove(cpp)
    print("built", target)


def main() -> int:
    if not os.path.exists(SRC):
        print("reference not present; keeping whatever is in oracle/_ref")
        return 0
    os.makedirs(OUT, exist_ok=True)
    _build_one(SRC, "ref_c_segmentation")
    # the ADAPTed LLR kernels: needed only to RUN the reference's adapter detection when real-read
    # golden fixtures are generated (oracle/make_golden_real.py); detection is upstream of the path
    llr = os.path.join(REF, "warpdemux", "adapted", "adapted", "detect", "_c_llr.pyx")
    if os.path.exists(llr):
        try:
            _build_one(llr, "ref_c_llr")
        except Exception as e:  # noqa: BLE001
            print("ref_c_llr not built:", e)
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""In-tree build of the CUDA library and the Cython host module (no network, no pip).

    python pyfastani_b200/build.py        # (a script on purpose: the package itself needs the built module)

1. make -C csrc          -> lib/libfastani_b200.so   (nvcc, -gencode arch=compute_100a,code=sm_100a)
2. cython _fastani.pyx   -> build/_fastani.cpp
3. g++ -shared           -> _fastani.<abi>.so, linked to the library with an $ORIGIN rpath
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _newer(src, dst):
    return not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(verbose=False):
    run = lambda cmd, **kw: subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL, **kw)
    jobs = str(min(8, os.cpu_count() or 1))
    run(["make", "-j", jobs, "-C", os.path.join(HERE, "csrc")])
    lib = os.path.join(HERE, "lib", "libfastani_b200.so")
    pyx = os.path.join(HERE, "_fastani.pyx")
    cpp = os.path.join(HERE, "build", "_fastani.cpp")
    ext = os.path.join(HERE, "_fastani" + sysconfig.get_config_var("EXT_SUFFIX"))
    header = os.path.join(ROOT, "include", "fastani_b200.h")
    os.makedirs(os.path.dirname(cpp), exist_ok=True)
    if _newer(pyx, cpp):
        run([sys.executable, "-m", "cython", "--cplus", "-3", pyx, "-o", cpp])
    if _newer(cpp, ext) or _newer(lib, ext) or _newer(header, ext):
        run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-w",
             "-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(ROOT, "include"),
             cpp, "-o", ext, "-L" + os.path.join(HERE, "lib"), "-lfastani_b200", "-Wl,-rpath,$ORIGIN/lib"])
    return lib, ext


if __name__ == "__main__":
    print("\n".join(build(verbose=True)))

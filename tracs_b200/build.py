"""Builds libtracs_b200.so (CUDA, sm_100a only) in-tree with nvcc. No JIT cache, no torch extension:
the built .so sits next to this file so it travels to the GPU box with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtracs_b200.so")
SOURCES = ["sweep.cu", "trans.cu", "capi.cu", "fasta.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return "nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f != "pybind_module.cpp"] + [os.path.join(HERE, "..", "include", "tracs_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    if not force and not stale():
        return LIB
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    for src in SOURCES:
        obj = os.path.join(bdir, src + ".o")
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)))
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== %s ==\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([_nvcc(), "-shared", "-o", LIB] + objs + ["-lz", "-lcudart"], env=env)
    return LIB


def pybind_path():
    import sysconfig
    return os.path.join(HERE, "dropin_native", "TRACS" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_pybind(force=False):
    """The compiled `TRACS` extension module (pybind11) over the C ABI: tracs_b200/dropin_native/TRACS<ext>.so, linked
    against libtracs_b200.so next door (rpath $ORIGIN/..)."""
    import sysconfig
    out = pybind_path()
    src = os.path.join(CSRC, "pybind_module.cpp")
    deps = [src, os.path.join(HERE, "..", "include", "tracs_b200.h"), LIB]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps if os.path.exists(d)):
        return out
    import pybind11
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-fvisibility=hidden", "-I" + pybind11.get_include(),
           "-I" + sysconfig.get_paths()["include"], src, "-o", out, "-L" + HERE, "-ltracs_b200", "-Wl,-rpath,$ORIGIN/.."]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_lib(force=True, verbose="-v" in sys.argv))
    print(build_pybind(force=True))

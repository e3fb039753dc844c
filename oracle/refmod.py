"""Loader for the UNMODIFIED reference module built into oracle/_ref/ by build_ref.sh.

TEST INFRASTRUCTURE ONLY. The module is named TRACS like the product's drop-in, so it is
loaded by explicit file path under a private name and never placed on sys.path."""
import importlib.machinery
import importlib.util
import os
import subprocess
import sys
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_MOD = None


def _path(variant):
    return os.path.join(_HERE, "_ref", variant, "TRACS" + sysconfig.get_config_var("EXT_SUFFIX"))


def available():
    return os.path.exists(_path("v3")) or os.path.exists(_path("native"))


def _runs(path):
    code = ("import importlib.util,importlib.machinery,sys;l=importlib.machinery.ExtensionFileLoader('TRACS',%r);"
            "s=importlib.util.spec_from_file_location('TRACS',%r,loader=l);m=importlib.util.module_from_spec(s);"
            "l.exec_module(m);m.trans_dist([1],[0.01],29.9,73.0,0.01)") % (path, path)
    return subprocess.run([sys.executable, "-c", code], capture_output=True).returncode == 0


def load():
    """Returns (module, variant). Prefers the -march=native build (the shipped flags); falls back
    to x86-64-v3 if the host CPU cannot run it."""
    global _MOD
    if _MOD is None:
        for variant in ("native", "v3"):
            p = _path(variant)
            if os.path.exists(p) and _runs(p):
                loader = importlib.machinery.ExtensionFileLoader("TRACS", p)
                spec = importlib.util.spec_from_file_location("TRACS", p, loader=loader)
                mod = importlib.util.module_from_spec(spec)
                loader.exec_module(mod)
                _MOD = (mod, variant)
                break
        else:
            raise RuntimeError("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    return _MOD

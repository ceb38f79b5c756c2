"""Builds tests/hostlib/lib*.so: host compilations of the device headers (see prim.cuh)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(name):
    src = os.path.join(HERE, name + ".cpp"); out = os.path.join(HERE, "lib" + name.replace("_", "") + ".so")
    csrc = os.path.join(HERE, "..", "..", "spartan2_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    return out

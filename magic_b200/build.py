"""Builds magic_b200/libmagic_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmagic_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-ccbin", "/usr/bin/g++"]


def sources():
    return [os.path.join(SRC, f) for f in sorted(os.listdir(SRC))] + [os.path.join(HERE, "..", "include", "magic_sht.h")]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False, out=None, defines=()):
    """out/defines: build an experimental variant (e.g. defines=["MAGIC_FFT_R_BIG=2"]) next to the default library."""
    global OUT
    if out is None:
        if not force and not needs_build():
            return OUT
        out = OUT
    cmd = [NVCC] + FLAGS + [f"-D{d}" for d in defines] + ["-o", out, os.path.join(SRC, "lib.cu"), "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libmagic_b200.so")
    with open(os.path.join(HERE, "ptxas_report.txt"), "w") as f:
        f.write(r.stdout)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose=True, out=outs[0] if outs else None, defines=defs))

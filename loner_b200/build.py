"""Builds the C-ABI shared library (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m loner_b200.build            # -> loner_b200/libloner_b200.so

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libloner_b200.so")
SOURCES = ["rays.cu", "sampler.cu", "render.cu", "mlp.cu", "hashgrid.cu", "optim.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "loner_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {s}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {s}")
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    if verbose:
        print("\n".join(log))
    return LIB


def build_probe():
    """Hardware probes used while sizing the kernels (tests/gpu_probe.py); NOT part of the product library."""
    d = os.path.join(HERE, "..", "tests", "probes")
    srcs = [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".cu")]
    out = os.path.join(d, "libloner_probe.so")
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call([NVCC] + FLAGS + ["-I", CSRC, "-shared"] + srcs + ["-o", out, "-lcudart"])
    return out


def build_trace():
    """mlp.cu with -DLONER_TRACE (timeline instrumentation of the pipelined kernels, tests/gpu_trace.py) as a
    separate probe library; NOT part of the product library."""
    d = os.path.join(HERE, "..", "tests", "probes")
    src = os.path.join(CSRC, "mlp.cu")
    out = os.path.join(d, "libloner_trace.so")
    if not os.path.exists(out) or os.path.getmtime(src) > os.path.getmtime(out):
        subprocess.check_call([NVCC] + [f for f in FLAGS if f not in ("-Xptxas", "-v")] + ["-DLONER_TRACE", "-shared", src, "-o", out, "-lcudart"])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

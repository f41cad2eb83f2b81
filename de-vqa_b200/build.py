"""Build libdevqa.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python de-vqa_b200/build.py [--force] [--verbose]

The .so lands next to this file so that it travels to the GPU box with the repo snapshot.
cudart is linked statically and NCCL is dlopen'ed at run time, so the library loads (and
exports every symbol of include/devqa.h) on machines with no GPU, driver or NCCL.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdevqa.so")
SOURCES = ["api.cu", "scan_stream.cu", "scan_tc.cu", "merge.cu", "refine.cu", "select.cu", "gate.cu", "vmm.cu", "synth.cu", "norm.cu", "nccl_dyn.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall", "-DDEVQA_BUILD",
         "-cudart", "static"]
if os.environ.get("DEVQA_TC_TRACE"):      # role timeline tracing of the tcgen05 scan (tools/tc_trace.py)
    FLAGS.append("-DDEVQA_TC_TRACE")


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "devqa.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.rsplit(".", 1)[0] + ".o")
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libdevqa.so")
    link = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
            "-Xcompiler", "-fPIC", "-Xlinker", "--exclude-libs,ALL", "-o", OUT, *objs, "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of libdevqa.so failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

"""Builds libmmnn_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmmnn_b200.so")
SOURCES = ["capi_conv.cu", "encoder.cu", "heads.cu", "optim.cu", "preprocess.cu", "augment.cu", "resnet.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--use_fast_math" if False else "-DMMNN_NO_FAST_MATH"]


def _newer(src_list, target):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("MMNN_EXTRA_NVCC_FLAGS", "").split()     # experiments only (e.g. -DMMNN_BRICK_TEST_ALIGNED=1)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]
    objs = [os.path.join(CSRC, os.path.splitext(os.path.basename(s))[0] + ".o") for s in srcs]

    def compile_one(pair):
        s, o = pair
        if force or _newer([s] + headers, o):
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
            if verbose:
                sys.stderr.write(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(compile_one, zip(srcs, objs)))
    if force or _newer(objs, OUT):
        # the CUDA runtime is linked DYNAMICALLY: the process already holds libcudart.so.12 (torch loads its own copy before this
        # library is opened), and a statically linked runtime would embed every runtime entry point's name -- including calls
        # this code base never makes -- as strings in the shipped binary
        cmd = [nvcc, "-shared", "--cudart", "shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                              "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""In-tree build of libhiercom_b200.so (nvcc, sm_100a only).  `python -m scene_graph_commonsense_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhiercom_b200.so")
SOURCES = ["abi.cu", "pairs.cu", "prep.cu", "blocks.cu", "head.cu", "topk.cu", "sgb.cu", "frontend.cu", "train.cu", "comm.cu", "tc_gemm.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--use_fast_math=false"]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "hiercom_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src, os.path.join(HERE, "..", "include", "hiercom_b200.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)
                                                                            if f.endswith((".cuh", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compiles the stale sources in parallel (one nvcc per file) and links; `force` recompiles everything."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    objs = [os.path.join(CSRC, src.replace(".cu", ".o")) for src in SOURCES]

    def compile_one(args):
        src, obj = args
        if not force and not _stale(os.path.join(CSRC, src), obj):
            return src, 0, "", ""
        r = subprocess.run([nvcc()] + flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        return src, r.returncode, r.stdout, r.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, zip(SOURCES, objs)))
    log = []
    for src, rc, out, err in results:
        if err:
            log.append("== %s\n%s" % (src, err))
        if rc != 0:
            sys.stderr.write(out + err)
            raise RuntimeError("nvcc failed on " + src)
    cmd = [nvcc(), "-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(CSRC, "ptxas.log"), "a" if not force else "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Build libgimb200.so in-tree with nvcc for sm_100a (no torch extension machinery: the library is a
plain C-ABI shared object, see include/gimb200.h)."""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgimb200.so")
TEST_LIB = os.path.join(HERE, "libgimb200_test.so")  # product objects + csrc/test_hooks.cu (include/gimb200_test.h)
SOURCES = ["common.cu", "engine.cu", "conv_simt.cu", "transformer.cu", "coarse_match.cu", "fine.cu", "umma_gemm.cu", "corr_sweep.cu", "loftr_api.cu", "dkm_kernels.cu", "dkm_api.cu"]
TEST_SOURCES = ["test_hooks.cu", "probe_mma.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isfile(c) or c == "nvcc"):
            return c
    raise RuntimeError("nvcc not found")


def _stale():
    if not os.path.isfile(LIB) or not os.path.isfile(TEST_LIB):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(TEST_LIB))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "gimb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, SOURCES + TEST_SOURCES))
    prod = objs[:len(SOURCES)]
    for lib, members in ((LIB, prod), (TEST_LIB, objs)):
        r = subprocess.run([nvcc, "-shared", "-o", lib, *members, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))

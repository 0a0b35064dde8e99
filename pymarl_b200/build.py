"""Build libpymarl_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension
machinery: the library is a plain C-ABI shared object loaded with ctypes)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(HERE, "libpymarl_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-Xptxas", "-v"] + os.environ.get("PMB_EXTRA_NVCC_FLAGS", "").split()


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_header_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link libpymarl_b200.so.  Incremental: an object is
    rebuilt when its source or any header is newer.  Returns the library path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_m = _newest_header_mtime()
    objs, rebuilt = [], False
    logs = []
    todo = []
    for src in _sources():
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ_DIR, src[:-3] + ".o")
        objs.append(op)
        if (not force and os.path.exists(op)
                and os.path.getmtime(op) >= max(os.path.getmtime(sp), hdr_m)):
            continue
        todo.append((src, [NVCC] + ARCH_FLAGS + COMMON + ["-c", sp, "-o", op]))
    if todo:
        # one nvcc per translation unit, in parallel (a clean build is ~1 minute serially)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(todo), max(1, (os.cpu_count() or 2)))) as pool:
            results = list(pool.map(lambda item: (item[0], item[1], subprocess.run(item[1], capture_output=True, text=True)), todo))
        for src, cmd, r in results:
            logs.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout + r.stderr))
        rebuilt = True
    if rebuilt or force or not os.path.exists(LIB_PATH):
        cmd = [NVCC] + ARCH_FLAGS + ["-shared", "-o", LIB_PATH] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(os.path.join(OBJ_DIR, "build.log"), "a" if not force else "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

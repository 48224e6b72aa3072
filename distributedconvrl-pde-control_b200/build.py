"""In-tree build of libpdeb200.so (sm_100a only) with plain nvcc.

`python distributedconvrl-pde-control_b200/build.py` or `build()`; objects are
cached under build/ by mtime, the library is written next to this file so that
it travels to the GPU box with the repository snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = ROOT / "build" / "obj"
LIB = PKG / "libpdeb200.so"
HOST_LIB = PKG / "libpdeb200_host.so"      # native host-side driver (C ABI only, g++; used by bench.py's e2e leg)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime():
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp")) + [ROOT / "include" / "pdeb200.h"]
    return max(p.stat().st_mtime for p in hdrs)


def _defines():
    return [x for x in os.environ.get("PDEB_EXTRA_DEFINES", "").split() if x]


def _compile(src, verbose):
    obj = OBJ / (src.stem + ".o")
    newest = max(src.stat().st_mtime, _deps_mtime(), Path(__file__).stat().st_mtime)
    if obj.exists() and obj.stat().st_mtime >= newest:
        return obj, False
    cmd = ["nvcc", *NVCC_FLAGS, *_defines(), "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src.name, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    return obj, True


def build(verbose=False, force=False):
    OBJ.mkdir(parents=True, exist_ok=True)
    if force:
        for o in OBJ.glob("*.o"):
            o.unlink()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), sources()))
    objs = [str(o) for o, _ in results]
    rebuilt = any(r for _, r in results)
    if rebuilt or not LIB.exists():
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *objs,
               "-Xcompiler", "-fPIC", "-ldl"]   # static cudart: self-contained next to torch's own runtime; NCCL is dlopen'ed
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    host_src = CSRC / "host" / "e2e_driver.cpp"
    if host_src.exists() and (rebuilt or not HOST_LIB.exists() or HOST_LIB.stat().st_mtime < host_src.stat().st_mtime):
        cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", str(host_src), "-o", str(HOST_LIB),
               "-L" + str(PKG), "-l:libpdeb200.so", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host driver build failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))

"""In-tree build of the native pieces (no JIT cache: the .so files travel with the repo snapshot).

Targets
  librnb       resnet_c_b200/librnb.so   — C-ABI (include/rnb.h): all sm_100a kernels + engine
  selftest     build/conv_selftest       — standalone GPU check of the tcgen05 conv kernel
  oracle       oracle/libref_ops.so      — plain-C restatement of the reference ops (gcc, CPU)
  ref          oracle/_ref/...           — the reference's own CUDA sources (only where /root/reference exists)
  dropin       build/resnet_infer, build/ref_main_dropin — drivers built on the cuda/*.cuh drop-in headers

`python -m resnet_c_b200.build [target ...]` builds the named targets (default: all that apply).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "resnet_c_b200" / "csrc"
LIB = ROOT / "resnet_c_b200" / "librnb.so"
REFERENCE = Path(os.environ.get("RNB_REFERENCE_DIR", "/root/reference"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]

LIB_SOURCES = ["api.cu", "model.cu", "conv_plan.cu", "tensormap.cu", "ops_f32.cu", "layout.cu",
               "stem.cu", "stem_tc.cu", "stem_tc_split.cu", "tail.cu", "group.cu", "preprocess.cu", "block.cu", "fp8.cu",
               "host_pack.cpp"]   # (.cpp: plain host code, compiled by g++ — AVX2 intrinsics, no CUDA)
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-pthread"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (no CPU fallback exists)")


def _run(cmd: list[str], cwd: Path = ROOT) -> None:
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], cwd=str(cwd), check=True)


def _stale(target: Path, sources: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources if s.exists())


def _includes(src: Path, seen: set[Path]) -> set[Path]:
    """Transitive quoted #includes of `src` inside the repo (dependency scan for incremental builds)."""
    import re
    if src in seen or not src.exists():
        return seen
    seen.add(src)
    for inc in re.findall(r'^\s*#\s*include\s*"([^"]+)"', src.read_text(), flags=re.M):
        _includes((src.parent / inc).resolve(), seen)
    return seen


def build_librnb(force: bool = False) -> Path:
    """One object per translation unit under build/obj/ (compiled in parallel, only when the source or one of the
    headers it includes changed), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    objdir = ROOT / "build" / "obj"
    objdir.mkdir(parents=True, exist_ok=True)
    jobs = []
    objs = []
    for name in LIB_SOURCES:
        src = CSRC / name
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, sorted(_includes(src.resolve(), set()))):
            if src.suffix == ".cpp":
                jobs.append([os.environ.get("CXX", "g++"), *CXX_FLAGS, "-c", src, "-o", obj])
            else:
                jobs.append([_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(_run, jobs))
    if jobs or force or _stale(LIB, objs):
        _run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs])
    return LIB


def build_selftest(force: bool = False) -> Path:
    out = ROOT / "build" / "conv_selftest"
    out.parent.mkdir(exist_ok=True)
    srcs = [ROOT / "tools" / "conv_selftest.cu", CSRC / "conv_plan.cu", CSRC / "tensormap.cu", CSRC / "fp8.cu"]
    deps = srcs + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh"))
    if force or _stale(out, deps):
        _run([_nvcc(), *NVCC_FLAGS, "-Xcompiler", "-fopenmp", "-o", out, *srcs])
    return out


def build_oracle(force: bool = False) -> Path:
    out = ROOT / "oracle" / "libref_ops.so"
    src = ROOT / "oracle" / "ref_ops.c"
    if force or _stale(out, [src]):
        _run(["make", "-C", ROOT / "oracle", "libref_ops.so"])
    return out


def build_ref(force: bool = False) -> Path | None:
    """Compile the reference's own CUDA sources where they lie (never copied into the repo)."""
    mk = ROOT / "oracle" / "Makefile"
    if not (REFERENCE / "cuda" / "ops.cu").exists() or not mk.exists():
        return None
    _run(["make", "-C", ROOT / "oracle", f"REFERENCE={REFERENCE}", "ref"])
    return ROOT / "oracle" / "_ref"


def build_dropin(force: bool = False) -> Path | None:
    mk = ROOT / "cuda" / "Makefile"
    if not mk.exists():
        return None
    build_librnb()
    _run(["make", "-C", ROOT / "cuda", f"REFERENCE={REFERENCE}"])
    return ROOT / "build"


TARGETS = {
    "librnb": build_librnb,
    "selftest": build_selftest,
    "oracle": build_oracle,
    "ref": build_ref,
    "dropin": build_dropin,
}


def build_all(force: bool = False) -> None:
    build_librnb(force)
    build_selftest(force)
    if (ROOT / "oracle" / "ref_ops.c").exists():
        build_oracle(force)
    build_ref(force)
    build_dropin(force)


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("-")]
    force = "--force" in sys.argv
    if not names:
        build_all(force)
    else:
        for n in names:
            TARGETS[n](force)

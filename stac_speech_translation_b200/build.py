"""In-tree build of libstac_b200.so (nvcc, sm_100a only; cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libstac_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", extra_flags=()) -> Path:
    """Default: libstac_b200.so.  `variant` + `extra_flags` build libstac_b200_<variant>.so next to it with additional
    nvcc flags (timing experiments, e.g. -DMHA_OSTAGED_PER_BUFFER); load it with STAC_B200_LIB=<path>."""
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [PKG.parent / "include" / "stac_b200.h"]
    objdir = CSRC / ("build_" + variant if variant else "build")
    objdir.mkdir(exist_ok=True)
    lib_path = PKG / f"libstac_b200_{variant}.so" if variant else LIB

    def compile_one(src: Path):
        obj = objdir / (src.stem + ".o")
        if force or _stale(obj, [src, *hdrs]):
            cmd = [NVCC, *FLAGS, *extra_flags, "-c", str(src), "-o", str(obj)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            (objdir / (src.stem + ".ptxas.log")).write_text(r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stderr}")
            if verbose:
                print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _stale(lib_path, objs):
        cmd = [NVCC, "-shared", "-o", str(lib_path), *map(str, objs), "-gencode",
               "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-lcuda"]
        # libcuda is resolved at run time through cudaGetDriverEntryPoint; no link-time dependency
        cmd = [c for c in cmd if c != "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr}")
    return lib_path


if __name__ == "__main__":
    # python -m stac_speech_translation_b200.build [--force] [-v] [--variant NAME -- <extra nvcc flags>]
    args = sys.argv[1:]
    extra = args[args.index("--") + 1:] if "--" in args else []
    name = args[args.index("--variant") + 1] if "--variant" in args else ""
    print(build(force="--force" in args, verbose="-v" in args, variant=name, extra_flags=extra))

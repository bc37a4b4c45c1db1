"""TEST INFRASTRUCTURE ONLY: builds tests/simt_cpu/libstac_simt_cpu.so, the library's plain SIMT kernels compiled from
their real source (csrc/*.cu) against tests/simt_cpu/cuda_runtime.h, a CPU emulation of the CUDA execution model.

Source rewriting is limited to what C++ cannot parse or link: `kernel<<<grid, block, smem, stream>>>(args);` becomes
`simt::launch(grid, block, smem, [&] { kernel(args); });` and `extern __shared__ T name[];` becomes a pointer to the
launch's dynamic shared-memory buffer.  Everything else - kernels, device helpers of common.cuh, the extern "C" entry
points with their argument checks - is compiled as written, so the CPU suite calls the same C ABI on host memory."""
import re
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE.parent.parent / "stac_speech_translation_b200" / "csrc"
SOURCES = ["turns.cu", "decoder_f32.cu", "ingest.cu", "norm_stats.cu", "encoder_f32.cu", "conv_frontend.cu", "fbank.cu",
           "train_pieces.cu"]
LIB = HERE / "libstac_simt_cpu.so"


def _split_top_level(text):
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite(src: str) -> str:
    out, pos = "", 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            return _dyn_smem(out + src[pos:])
        j = src.index(">>>", i)
        name_start = max(src.rfind(c, 0, i) for c in " \n\t;{}") + 1
        name = src[name_start:i]
        cfg = _split_top_level(src[i + 3:j])
        k = j + 3
        assert src[k] == "(", "launch arguments expected"
        depth, e = 0, k
        while True:
            depth += src[e] == "("
            depth -= src[e] == ")"
            if depth == 0:
                break
            e += 1
        args = src[k + 1:e]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += src[pos:name_start] + f"simt::launch(dim3({cfg[0]}), dim3({cfg[1]}), {smem}, [&] {{ {name}({args}); }})"
        pos = e + 1


def _dyn_smem(src: str) -> str:
    return re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?((?:unsigned\s+)?\w+)\s+(\w+)\[\];",
                  r"\1* \2 = reinterpret_cast<\1*>(simt::dyn_smem);", src)


def build(force: bool = False) -> Path:
    deps = [CSRC / s for s in SOURCES] + [CSRC / "common.cuh", CSRC / "gemm_simt.cuh", HERE / "cuda_runtime.h",
                                          Path(__file__)]
    if LIB.exists() and not force and all(d.stat().st_mtime < LIB.stat().st_mtime for d in deps):
        return LIB
    gen = HERE / "_generated"
    gen.mkdir(exist_ok=True)
    units = []
    for s in SOURCES:
        text = rewrite((CSRC / s).read_text()).replace('#include "common.cuh"', f'#include "{CSRC / "common.cuh"}"')
        unit = gen / (Path(s).stem + ".cpp")
        unit.write_text(text)
        units.append(str(unit))
    (gen / "api_stub.cpp").write_text(
        '#include "cuda_runtime.h"\nint stac_grid_limit() { return 148; }\n'
        "// the tensor-core entry points the SIMT files hand over to are not part of this build\n"
        "int stac_conv0_tc_launch(const float*, const float*, const float*, const float*, const float*, long, long, int,\n"
        "                         unsigned short*, void*, const unsigned int*, int, float, const float*, const float*)\n"
        "{ return -2; }\n")
    cmd = ["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread", "-fpermissive", "-w", f"-I{HERE}", f"-I{CSRC}", *units,
           str(gen / "api_stub.cpp"), "-o", str(LIB)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("simt_cpu build failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))

"""Builds VARIANTS of libstx_b200.so for A/B timing on the GPU box (STX_B200_LIB=<variant> python tools/time_kernels.py K).

    python tools/ab_build.py name=-DMACRO[=value][,-DOTHER] [name2=...]      # e.g.  wide=-DSTX_EXPERIMENT_WIDE=1

Every variant compiles a private copy of csrc/ with the extra nvcc flags into build/variants/lib_<name>.so (build/ is
git-ignored but travels with `gpurun`; delete it afterwards: every variant adds ~3 MB to each snapshot).  The sources are
copied to a directory at the same depth as csrc/ so that their relative includes keep working.
"""
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "speech_transcript_embeddings_b200" / "csrc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def build_variant(name: str, extra: list[str]) -> Path:
    src_dir = ROOT / "build" / f"vsrc_{name}"                 # two levels below the root, like csrc/
    if src_dir.exists():
        shutil.rmtree(src_dir)
    shutil.copytree(CSRC, src_dir, ignore=shutil.ignore_patterns("_obj"))

    def cc(src: Path) -> str:
        obj = src_dir / (src.stem + ".o")
        subprocess.check_call(["nvcc", *FLAGS, *extra, "-c", str(src), "-o", str(obj)])
        return str(obj)

    with ThreadPoolExecutor(4) as ex:
        objs = list(ex.map(cc, sorted(src_dir.glob("*.cu"))))
    out_dir = ROOT / "build" / "variants"
    out_dir.mkdir(parents=True, exist_ok=True)
    out = out_dir / f"lib_{name}.so"
    subprocess.check_call(["nvcc", "-shared", "-o", str(out), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                           "-Xcompiler", "-fPIC"])
    shutil.rmtree(src_dir)
    return out


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition("=")
        print(build_variant(name, [f for f in flags.split(",") if f]))

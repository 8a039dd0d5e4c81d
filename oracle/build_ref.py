"""Build the UNMODIFIED reference (joeyan/gaussian_splatting) into oracle/_ref/ as the GPU oracle.

TEST INFRASTRUCTURE ONLY.  Nothing under gaussian_splatting_b200/ imports or links this.

What it does (only when /root/reference exists, i.e. in the build container; the GPU box uses the
prebuilt, git-ignored oracle/_ref/ that travels with the tree):

  * compiles the reference's src/bindings.cpp + 7 .cu files where they lie, with our own nvcc/g++
    command lines (the reference's setup.py is not run), for sm_100 and with the flags
    torch.utils.cpp_extension would have used for it (no -O3/-use_fast_math: setup.py:23-26 passes
    its flags to the wrong argument), as a module named `splat_cuda_ref`;
  * installs (copies) the reference's pure-python package `splat_py` and its `test/` directory next
    to it, exactly like a `pip install --target oracle/_ref` would.  No reference source enters the
    git history: oracle/_ref/ is listed in .gitignore.

Usage:  python oracle/build_ref.py [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
TMP = HERE / "_build" / "ref_obj"
REF = Path(os.environ.get("GSR_REFERENCE_DIR", "/root/reference"))

SOURCES = [
    "src/bindings.cpp",
    "src/depth.cu",
    "src/precompute_sh.cu",
    "src/projection.cu",
    "src/projection_backward.cu",
    "src/render.cu",
    "src/render_backward.cu",
    "src/tile_culling.cu",
]


def ext_path() -> Path:
    return OUT / f"splat_cuda_ref{sysconfig.get_config_var('EXT_SUFFIX')}"


def available() -> bool:
    return ext_path().exists() and (OUT / "splat_py" / "rasterize.py").exists()


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref is usable afterwards."""
    if available() and not force:
        return True
    if not REF.exists():
        return available()
    import torch
    from torch.utils import cpp_extension as ce

    OUT.mkdir(parents=True, exist_ok=True)
    TMP.mkdir(parents=True, exist_ok=True)
    incs = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    common = [
        "-DTORCH_EXTENSION_NAME=splat_cuda_ref", "-DTORCH_API_INCLUDE_EXTENSION_H",
        f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", *incs,
    ]
    nvcc = "/usr/local/cuda/bin/nvcc"

    def compile_one(rel):
        src = REF / rel
        obj = TMP / (Path(rel).stem + ".o")
        if rel.endswith(".cu"):
            cmd = [nvcc, "-c", str(src), "-o", str(obj), *common,
                   "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                   "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                   "--expt-relaxed-constexpr", "-gencode=arch=compute_100,code=sm_100",
                   "--compiler-options", "-fPIC", "-std=c++17"]
        else:
            cmd = ["g++", "-c", str(src), "-o", str(obj), *common, "-fPIC", "-std=c++17", "-O2"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError(f"reference compile failed: {rel}")
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    torch_lib = Path(torch.__file__).parent / "lib"
    cuda_lib = Path(ce.CUDA_HOME or "/usr/local/cuda") / "lib64"
    link = ["g++", "-shared", "-o", str(ext_path()), *map(str, objs), f"-L{torch_lib}", "-ltorch", "-ltorch_cpu",
            "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python", f"-L{cuda_lib}", "-lcudart",
            f"-Wl,-rpath,{torch_lib}", f"-Wl,-rpath,{cuda_lib}"]
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("reference link failed")
    for sub in ("splat_py", "test"):
        dst = OUT / sub
        if dst.exists():
            shutil.rmtree(dst)
        shutil.copytree(REF / sub, dst, ignore=shutil.ignore_patterns("__pycache__"))
        for p in [dst, *dst.rglob("*")]:  # the source tree is read-only; the installed copy must not be
            p.chmod(p.stat().st_mode | 0o200)
    (OUT / "README").write_text(
        "Installed copy of the unmodified reference (joeyan/gaussian_splatting) built by oracle/build_ref.py.\n"
        "Git-ignored; test oracle only.\n")
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference and no prebuilt copy)")

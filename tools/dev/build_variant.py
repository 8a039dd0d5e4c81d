"""Build a VARIANT of libgsr_b200.so (extra -D flags) into _dbg/variants/<name>/libgsr_b200.so.

    python tools/dev/build_variant.py stats -DGSR_STATS
    python tools/dev/build_variant.py bwd_mb7 -DGSR_BWD_MINB=7

Development aid for A/B timing on the GPU box (tools/dev/run_variants.sh swaps each variant in and runs
_dbg/time_stages.py); not part of the product build."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from gaussian_splatting_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out = ROOT / "_dbg" / "variants" / name
    out.mkdir(parents=True, exist_ok=True)
    objs = []
    procs = []
    for src in B.CU_SOURCES:
        obj = out / (Path(src).stem + ".o")
        cmd = [B._nvcc(), *B.NVCC_FLAGS, *flags, "-c", str(B.CSRC / src), "-o", str(obj)]
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), src))
        objs.append(obj)
    for p, src in procs:
        o, _ = p.communicate()
        if p.returncode != 0:
            sys.exit(f"{src}: {o}")
    subprocess.run([B._nvcc(), "-shared", "-o", str(out / "libgsr_b200.so"), *map(str, objs), "-gencode",
                    "arch=compute_100a,code=sm_100a", "-cudart", "shared"], check=True)
    for o in objs:
        o.unlink()
    print(out / "libgsr_b200.so")


if __name__ == "__main__":
    main()

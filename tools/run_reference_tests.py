"""Run the reference's OWN unittest files (installed copy under oracle/_ref/test) with its python
package bound to either CUDA implementation:

    python tools/run_reference_tests.py --impl b200   # reference splat_py + tests on THIS library (drop-in proof)
    python tools/run_reference_tests.py --impl ref    # same tests on the compiled reference (sanity / stale tests)

test_dataloader.py is skipped (it needs a dataset on the reference author's disk).
Writes a JSON summary to --out.  Needs a GPU.
"""
from __future__ import annotations

import argparse
import io
import json
import sys
import unittest
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF_DIR = ROOT / "oracle" / "_ref"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["b200", "ref"], required=True)
    ap.add_argument("--out", default=None)
    ap.add_argument("--module", default=None, help="run one test module in-process (used by the driver mode)")
    ap.add_argument("--per-module-timeout", type=int, default=120)
    ap.add_argument("--modules", default=None, help="comma-separated subset of test modules")
    ap.add_argument("--same-as", default=None,
                    help="JSON summary of another --impl run: exit 0 iff this run fails EXACTLY the same tests (the "
                         "reference's own stale literals fail on the reference build too)")
    args = ap.parse_args()
    names = ["test_projection", "test_tile_culling", "test_rasterize", "test_depth", "test_structs", "test_utils",
             "test_cuda_autograd_functions", "test_rasterize_autograd"]
    if args.modules:
        names = args.modules.split(",")
    if args.module is None:
        # driver mode: one subprocess per module, each with its own timeout, so a hang is isolated and named
        import subprocess
        import time

        summary = dict(impl=args.impl, modules={})
        for n in names:
            t0 = time.time()
            try:
                pr = subprocess.run([sys.executable, __file__, "--impl", args.impl, "--module", n], capture_output=True,
                                    text=True, timeout=args.per_module_timeout)
                tail = (pr.stdout + pr.stderr)[-1500:]
                status = "ok" if pr.returncode == 0 else f"rc={pr.returncode}"
                detail = None
                for line in reversed(pr.stdout.splitlines()):  # the child's last JSON line names the failing tests
                    if line.startswith("{") and '"failures"' in line:
                        detail = json.loads(line)
                        break
            except subprocess.TimeoutExpired as e:
                tail = ((e.stdout or b"").decode(errors="ignore") + (e.stderr or b"").decode(errors="ignore"))[-1500:]
                status = f"TIMEOUT after {args.per_module_timeout}s"
                detail = None
            summary["modules"][n] = dict(status=status, seconds=round(time.time() - t0, 1), tail=tail,
                                         run=None if detail is None else detail["run"],
                                         failed=None if detail is None else sorted(detail["failures"] + detail["errors"]))
            print(n, status, f"{time.time() - t0:.1f}s", flush=True)
            if args.out:
                Path(args.out).parent.mkdir(parents=True, exist_ok=True)
                Path(args.out).write_text(json.dumps(summary, indent=1))
        summary["ok"] = all(m["status"] == "ok" for m in summary["modules"].values())
        print(json.dumps({k: v["status"] for k, v in summary["modules"].items()}))
        def failing(sm):
            return {n: (m.get("failed") if m["status"].startswith("rc=") else m["status"])
                    for n, m in sm["modules"].items() if m["status"] != "ok"}

        rc = 0 if summary["ok"] else 1
        if args.same_as:
            other = json.loads(Path(args.same_as).read_text())
            summary["same_failures_as"] = dict(file=args.same_as, impl=other.get("impl"),
                                               identical=failing(other) == failing(summary), mine=failing(summary),
                                               theirs=failing(other))
            rc = 0 if summary["same_failures_as"]["identical"] else 1
            print(json.dumps(summary["same_failures_as"]))
        if args.out:
            Path(args.out).write_text(json.dumps(summary, indent=1))
        return rc
    names = [args.module]
    import torch  # noqa: F401

    if args.impl == "b200":
        import gaussian_splatting_b200 as g

        g.install_as_splat_cuda()
    else:
        from oracle import ref_loader

        sys.modules["splat_cuda"] = ref_loader._load_ref_ext()
    sys.path.insert(0, str(REF_DIR))          # reference `splat_py`
    sys.path.insert(0, str(REF_DIR / "test"))  # its fixtures module
    suite = unittest.TestSuite()
    loader = unittest.TestLoader()
    for n in names:
        suite.addTests(loader.loadTestsFromName(n))
    buf = io.StringIO()
    res = unittest.TextTestRunner(stream=buf, verbosity=2).run(suite)
    text = buf.getvalue()
    print(text[-6000:])
    summary = dict(impl=args.impl, run=res.testsRun, failures=[str(t[0]) for t in res.failures],
                   errors=[str(t[0]) for t in res.errors], ok=res.wasSuccessful(),
                   details={str(t[0]): t[1][-800:] for t in res.failures + res.errors})
    print(json.dumps({k: summary[k] for k in ("impl", "run", "failures", "errors", "ok")}))
    return 0 if res.wasSuccessful() else 1


if __name__ == "__main__":
    sys.exit(main())

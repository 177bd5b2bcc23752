#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's own implementation of the hot path, verbatim.

    python oracle/make_ref.py            # in the build container (needs /root/reference)

The reference is a pure-Python scratch repository with no build system, so "building"
it means copying the two modules that define the path -- qmc/quantization_model.py (linear domain) and
qmc/quantization_model_log.py (log domain) -- byte for byte into oracle/_ref/.  That directory is git-ignored
(reference sources never enter this repository's history) but travels to the GPU box with the repo snapshot,
where /root/reference does not exist.  bench.py's `--impl reference` arm and its `cpu_baseline` leg import the
copy (kind "reference"); without it they fall back to the oracle's port (kind "port").  Test infrastructure
only: nothing under quantized_spectrum_cartography_b200/ may import oracle/.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QMC_REFERENCE_ROOT", "/root/reference")
FILES = ("qmc/quantization_model.py", "qmc/quantization_model_log.py")


def main() -> int:
    if not os.path.isdir(REF):
        print(f"{REF} not present: keeping whatever oracle/_ref/ holds")
        return 0
    dst = os.path.join(HERE, "_ref")
    os.makedirs(dst, exist_ok=True)
    with open(os.path.join(dst, "MANIFEST.txt"), "w") as man:
        for rel in FILES:
            src = os.path.join(REF, rel)
            out = os.path.join(dst, os.path.basename(rel))
            shutil.copyfile(src, out)
            digest = hashlib.sha256(open(out, "rb").read()).hexdigest()
            man.write(f"{rel} sha256={digest}\n")
            print(f"copied {rel} -> oracle/_ref/{os.path.basename(rel)} ({digest[:16]})")
    return 0


if __name__ == "__main__":
    sys.exit(main())

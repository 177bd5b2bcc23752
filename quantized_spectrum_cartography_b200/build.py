"""Build libqmc_b200.so in-tree with nvcc for sm_100a.

    python quantized_spectrum_cartography_b200/build.py        # rebuild if sources are newer
    python quantized_spectrum_cartography_b200/build.py -f     # force
(run it as a script: importing the package needs the library to exist already)

The library is a plain CUDA shared object (C ABI, include/qmc_b200.h); it does not link against
torch.  The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libqmc_b200.so")
# (source, object stem, extra flags, headers it depends on besides the common ones)
FAMILIES = {1: "flat", 2: "tiled", 3: "lanes"}
RANKS = (1, 2, 4, 8, 16, 32)
COMMON_DEPS = ["qmc_common.cuh", "erfcx_coeffs.h", os.path.join("..", "..", "include", "qmc_b200.h")]


def units() -> list[tuple[str, str, list[str], list[str]]]:
    u = [("qmc_abi.cu", "qmc_abi", [], []), ("qmc_quantize.cu", "qmc_quantize", [], []),
         ("qmc_dense.cu", "qmc_dense", [], []), ("qmc_solver.cu", "qmc_solver", [], []), ("qmc_gather.cu", "qmc_gather", [], ["qmc_gather_common.cuh"]),
         ("qmc_lanes_build.cu", "qmc_lanes_build", [], ["qmc_gather_common.cuh"])]
    for fam, name in FAMILIES.items():
        for rp in RANKS:
            u.append(("qmc_gather_inst.cu", f"qmc_gather_{name}_r{rp}", [f"-DQMC_FAMILY={fam}", f"-DQMC_RP={rp}"],
                      ["qmc_gather_common.cuh", f"qmc_gather_{name}.cuh"]))
    return u


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _obj(stem: str) -> str:
    return os.path.join(PKG, "build", stem + ".o")


def _stale(stem: str, src: str, deps: list[str]) -> bool:
    obj = _obj(stem)
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    files = [os.path.join(CSRC, src)] + [os.path.join(CSRC, d) for d in COMMON_DEPS + deps]
    return any(os.path.getmtime(f) > t for f in files)


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(_stale(stem, src, deps) or os.path.getmtime(_obj(stem)) > t for src, stem, _, deps in units())


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> str:
    """Compile the stale translation units (``jobs`` at a time, default: all cores) and link."""
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    todo = [(src, stem, fl) for src, stem, fl, deps in units() if force or _stale(stem, src, deps)]
    jobs = jobs or int(os.environ.get("QMC_BUILD_JOBS", os.cpu_count() or 4))
    running: list[tuple[str, subprocess.Popen]] = []
    logs = {}

    def reap(block: bool):
        for item in list(running):
            stem, p = item
            if block or p.poll() is not None:
                out, _ = p.communicate()
                running.remove(item)
                logs[stem] = out
                if p.returncode != 0:
                    for _, q in running:
                        q.kill()
                    raise RuntimeError(f"nvcc failed on {stem}:\n{out}")
                if block:
                    return

    for src, stem, fl in todo:
        while len(running) >= jobs:
            reap(block=False)
            if len(running) >= jobs:
                import time
                time.sleep(0.2)
        cmd = [_nvcc(), *NVCC_FLAGS, *fl, "-c", os.path.join(CSRC, src), "-o", _obj(stem)]
        running.append((stem, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    while running:
        reap(block=True)
    objs = [_obj(stem) for _, stem, _, _ in units()]
    link = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    for stem, out in logs.items():
        with open(os.path.join(PKG, "build", stem + ".ptxas.log"), "w") as f:
            f.write(out)
    if verbose:
        print("\n".join(logs.values()))
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))

#!/usr/bin/env python
"""Count the SASS mnemonics that prove the Blackwell-native paths, per built object (cuobjdump -sass):
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk, TMA),
SYNCS (mbarrier), FFMA2/FMUL2/FADD2 (packed fp32x2), LDGSTS (cp.async), MUFU (SFU), REDUX/MATCH (builder).
    python tools/sass_markers.py > profiles/r2_sass_markers.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARK = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "LDGSTS",
        "MUFU", "REDUX", "MATCH", "HMMA", "ATOMS", "RED."]


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "quantized_spectrum_cartography_b200", "build", "*.o")))
    pick = [o for o in objs if re.search(r"qmc_(dense|lanes_build|solver|quantize|abi|gather_lanes_r4|gather_lanes_r16|gather_tiled_r4|gather_flat_r4)\.o$", o)]
    print("# cuobjdump -sass mnemonic counts per object (all kernels of the object); sm_100a, nvcc 12.9")
    print("# object".ljust(28) + "".join(m.rjust(9) for m in MARK))
    for o in pick:
        out = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        c = collections.Counter()
        for line in out.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1)
            for k in MARK:
                if op.startswith(k):
                    c[k] += 1
        print(os.path.basename(o).ljust(28) + "".join(str(c[k]).rjust(9) for k in MARK))
    # the headline kernel on its own
    o = os.path.join(ROOT, "quantized_spectrum_cartography_b200", "build", "qmc_gather_lanes_r4.o")
    fun = "_ZN3qmc19gather_lanes_kernelILi4ELi2ELb0ELi1ELb1EEEvNS_12GatherParamsE"
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, o], capture_output=True, text=True).stdout
    c = collections.Counter()
    n = 0
    for line in out.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            n += 1
            for k in MARK:
                if m.group(1).startswith(k):
                    c[k] += 1
    print(f"\n# gather_lanes_kernel<4, ONEBIT, linear, both gradients, 16-bit words>: {n} SASS instructions")
    print("  " + ", ".join(f"{k} {c[k]}" for k in MARK if c[k]))
    out = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3qmc12dense_kernelILi0ELb1ELb1ELi4EEEvNS_11DenseParamsE",
                          os.path.join(ROOT, "quantized_spectrum_cartography_b200", "build", "qmc_dense.o")], capture_output=True, text=True).stdout
    c = collections.Counter()
    n = 0
    for line in out.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            n += 1
            for k in MARK:
                if m.group(1).startswith(k):
                    c[k] += 1
    print(f"\n# dense_kernel<STABLE, log domain, gradients, 4 evaluations in flight> (cfg4): {n} SASS instructions")
    print("  " + ", ".join(f"{k} {c[k]}" for k in MARK if c[k]))


if __name__ == "__main__":
    main()

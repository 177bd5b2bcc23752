#!/usr/bin/env python
"""Lane-stream layout at cfg3: builder time, padding, steps per stream, and the shared-memory wavefronts per
S / gS row access predicted by the measured bank model (tools/micro/bank128*.cu: per quarter-warp, the largest
number of distinct 16-byte rows sharing a bank group)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench


def main():
    n_maps = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dev = torch.device("cuda", 0)
    wl = bench.build_workload(n_maps, dev, seed=0)
    obs = wl["obs"]
    from quantized_spectrum_cartography_b200.obs import ObsSet, lane_streams
    rows = ObsSet(obs.idx, obs.lvl, obs.row_off, obs.B, obs.K, obs.IJ, obs.n_sub, obs.sub_pixels, obs.tile_warps,
                  obs.nobs, obs.max_level)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o2 = lane_streams(rows)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    nr = obs.nrows.cpu().numpy()
    wf, acc = 0, 0
    for s in range(0, min(obs.B * obs.n_sub, 256)):
        lv, band, pix, real, row = obs.decode_stream(s)
        for t in range(pix.shape[0]):
            for qd in range(4):
                sl = slice(8 * qd, 8 * qd + 8)
                p = np.unique(pix[t, sl])        # padding shares a real lane's row
                if len(p):
                    wf += np.bincount(p & 7, minlength=8).max()
            acc += 1
    print(json.dumps({"maps": n_maps, "builder_ms": 1e3 * t_build, "padding": obs.padding_fraction(),
                      "steps_mean": float(nr.mean()), "steps_hist": {int(k): int(v) for k, v in zip(*np.unique(nr, return_counts=True))},
                      "wavefronts_per_access": wf / acc, "word_bits": obs.word_bits, "n_runs": obs.n_runs,
                      "stride_words": obs.stream_stride}))


if __name__ == "__main__":
    main()

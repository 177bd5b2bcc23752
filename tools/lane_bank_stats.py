"""Bank-group statistics of a lane-stream observation set (cfg3 geometry): expected LDS.128
wavefronts per step under the quarter-warp model, padding, band switches per group."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
import quantized_spectrum_cartography_b200 as q

wl = bench.build_workload(64, torch.device("cuda"), seed=0)
obs = wl["obs"]
words = obs.words.cpu().numpy().view(np.uint32).astype(np.int64)
so, nr = obs.stream_off.cpu().numpy(), obs.nrows.cpu().numpy()
wf_q, wf_any, steps, pads, sw_groups, groups = 0, 0, 0, 0, 0, 0
for s in range(len(nr)):
    blk = words[so[s]: so[s] + 32 * nr[s]].reshape(-1, 32, 4).transpose(0, 2, 1).reshape(-1, 32)
    res = (blk & 0x7FFF) % 8
    band = (blk >> 15) & 0x1FF
    real = (blk >> 24) != 0xFF
    res = np.where(real, res, -1 - np.arange(32)[None, :] // 8 * 0)
    for row in res:
        # padding lanes broadcast a real lane's row of their quarter: they add no wavefront
        wq = sum(max(1, np.bincount(row[8 * i: 8 * i + 8][row[8 * i: 8 * i + 8] >= 0], minlength=8).max()) for i in range(4))
        wf_q += wq
        wf_any += max(4, np.bincount(row[row >= 0], minlength=8).max())
    steps += len(blk)
    pads += (~real).sum()
    g = band.reshape(-1, 4, 32)
    sw_groups += (g[:, 0, :] != g[:, 3, :]).any(axis=1).sum() + (g[1:, 0, :] != g[:-1, 3, :]).any(axis=1).sum() * 0
    groups += len(g)
print(f"steps {steps}  padding {pads / (steps * 32):.4f}")
print(f"wavefronts/step quarter-warp model {wf_q / steps:.3f}  (ideal 4)   any-8-lanes model {wf_any / steps:.3f}")
print(f"groups with a band switch inside {sw_groups / groups:.3f}")
import os
os.makedirs("gpurun_out", exist_ok=True)
n_keep = 16 * obs.n_sub
np.savez_compressed("gpurun_out/lane_words.npz", words=words[: so[n_keep]].astype(np.uint32), so=so[: n_keep + 1], nr=nr[:n_keep],
                    n_sub=obs.n_sub, sub=obs.sub_pixels, tw=obs.tile_warps)

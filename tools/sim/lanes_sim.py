#!/usr/bin/env python
"""CPU model of the lane-stream builder (qmc_obs_build_lanes): step count, padding and shared-memory
wavefronts per access for layout variants.  Development tool (no GPU needed)."""
import sys
import numpy as np


def make_stream(rng, K=64, sub_pixels=326, f=0.10, pix0=0):
    """entries[k] = array of tile-local pixels observed in band k"""
    m = rng.random((K, sub_pixels)) < f
    return [np.nonzero(m[k])[0] + pix0 for k in range(K)]


def max_matching(cands):
    """cands[i] = set of classes lane i may take.  Returns dict lane->class (maximum matching, lanes in order,
    best-stocked class first is approximated by the caller's ordering of cands[i])."""
    owner = {}

    def aug(i, seen):
        for c in cands[i]:
            if c in seen:
                continue
            seen.add(c)
            if c not in owner or aug(owner[c], seen):
                owner[c] = i
                return True
        return False
    for i in range(len(cands)):
        if cands[i]:
            aug(i, set())
    return {i: c for c, i in owner.items()}


def wavefronts(pixels):
    """LDS.128 wavefronts of one access: per quarter-warp the largest number of distinct rows per bank group."""
    tot = 0
    for q in range(4):
        rows = {}
        for p in pixels[8 * q: 8 * q + 8]:
            if p is None:
                continue
            rows.setdefault(p & 7, set()).add(p)
        tot += max((len(v) for v in rows.values()), default=0)
    return tot


def build(entries, mode="snake", group_distinct=False, regroup=True):
    """Greedy step-by-step builder.  Returns (steps, wavefronts per step list, n_entries)."""
    K = len(entries)
    sizes = np.array([len(e) for e in entries])
    order = sorted(range(K), key=lambda k: (-sizes[k], k))
    G = (K + 31) // 32
    lanes = [[] for _ in range(32)]   # list of bands per lane (static modes)
    if mode == "snake":
        for g in range(G):
            for l in range(32):
                r = 32 * g + (31 - l if g & 1 else l)
                if r < K:
                    lanes[l].append(order[r])
    pools = {k: list(entries[k]) for k in range(K)}
    queue = None
    if mode == "dynamic":            # bands taken from a queue, largest first; lanes pull when idle
        queue = list(order)
    cur = [None] * 32                 # current band per lane
    steps = 0
    wf = []
    group_used = set()
    nslots_pad = 0
    total = int(sizes.sum())
    left = total
    while left > 0:
        if steps % 4 == 0:
            group_used = set()
            for l in range(32):
                if cur[l] is None or not pools[cur[l]]:
                    cur[l] = None
                    if queue is not None:
                        while queue and not pools[queue[0]]:
                            queue.pop(0)
                        if queue:
                            cur[l] = queue.pop(0)
                    else:
                        while lanes[l] and not pools[lanes[l][0]]:
                            lanes[l].pop(0)
                        if lanes[l]:
                            cur[l] = lanes[l].pop(0)
        chosen = [None] * 32
        taken = set(group_used) if group_distinct else set()
        for q in range(4):
            ls = list(range(8 * q, 8 * q + 8))
            cands = []
            for l in ls:
                if cur[l] is None:
                    cands.append([])
                    continue
                cnt = {}
                for p in pools[cur[l]]:
                    if p in taken:
                        continue
                    cnt[p & 7] = cnt.get(p & 7, 0) + 1
                cands.append(sorted(cnt, key=lambda c: -cnt[c]))
            m = max_matching(cands)
            for i, l in enumerate(ls):
                if cur[l] is None:
                    continue
                pool = pools[cur[l]]
                pick = None
                if i in m:
                    for p in pool:
                        if (p & 7) == m[i] and p not in taken:
                            pick = p
                            break
                if pick is None:
                    for p in pool:
                        if p not in taken:
                            pick = p
                            break
                if pick is not None:
                    chosen[l] = pick
                    taken.add(pick)
                    pool.remove(pick)
                    left -= 1
        group_used |= {p for p in chosen if p is not None}
        wf.append(wavefronts(chosen))
        nslots_pad += sum(1 for p in chosen if p is None)
        steps += 1
        if steps > 4000:
            raise RuntimeError("no progress")
    while steps % 4:
        steps += 1
        wf.append(0)
    return steps, wf, total


def main():
    rng = np.random.default_rng(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    for mode, gd in (("snake", False), ("snake", True), ("dynamic", False), ("dynamic", True)):
        st, wfs, ents = [], [], []
        for i in range(n):
            e = make_stream(np.random.default_rng(i))
            s, wf, t = build(e, mode=mode, group_distinct=gd)
            st.append(s)
            ents.append(t)
            wfs.append(sum(wf) / max(1, sum(1 for w in wf if w)))
        print(f"{mode:8s} group_distinct={gd}: steps {np.mean(st):.2f} (ideal {np.mean(ents) / 32:.2f}), padding "
              f"{1 - np.sum(ents) / (32 * np.sum(st)):.4f}, wavefronts/access {np.mean(wfs):.3f}")


if __name__ == "__main__":
    main()


def analyse(n=8):
    """Where do the extra wavefronts come from?  Histogram of (min pool size in the quarter) vs conflicts."""
    import collections
    tot = collections.Counter()
    cnt = collections.Counter()
    for i in range(n):
        e = make_stream(np.random.default_rng(i))
        K = len(e)
        sizes = np.array([len(x) for x in e])
        order = sorted(range(K), key=lambda k: (-sizes[k], k))
        lanes = [[] for _ in range(32)]
        for g in range(2):
            for l in range(32):
                r = 32 * g + (31 - l if g & 1 else l)
                lanes[l].append(order[r])
        pools = {k: list(e[k]) for k in range(K)}
        cur = [None] * 32
        left = int(sizes.sum())
        steps = 0
        while left > 0:
            if steps % 4 == 0:
                for l in range(32):
                    if cur[l] is None or not pools[cur[l]]:
                        cur[l] = lanes[l].pop(0) if lanes[l] else None
            taken = set()
            for q in range(4):
                ls = list(range(8 * q, 8 * q + 8))
                cands = []
                minpool = 99
                for l in ls:
                    if cur[l] is None:
                        cands.append([])
                        continue
                    cntc = {}
                    for p in pools[cur[l]]:
                        if p not in taken:
                            cntc[p & 7] = cntc.get(p & 7, 0) + 1
                    minpool = min(minpool, len(pools[cur[l]]))
                    cands.append(sorted(cntc, key=lambda c: -cntc[c]))
                m = max_matching(cands)
                chosen = []
                for i2, l in enumerate(ls):
                    if cur[l] is None:
                        continue
                    pool = pools[cur[l]]
                    pick = None
                    if i2 in m:
                        for p in pool:
                            if (p & 7) == m[i2] and p not in taken:
                                pick = p
                                break
                    if pick is None:
                        for p in pool:
                            if p not in taken:
                                pick = p
                                break
                    if pick is not None:
                        chosen.append(pick)
                        taken.add(pick)
                        pool.remove(pick)
                        left -= 1
                rows = {}
                for p in chosen:
                    rows.setdefault(p & 7, set()).add(p)
                w = max((len(v) for v in rows.values()), default=0)
                b = min(minpool, 12)
                tot[b] += w
                cnt[b] += 1
            steps += 1
    for b in sorted(cnt):
        print(f"min pool in quarter {b:2d}: accesses {cnt[b]:5d}  mean wavefronts/quarter {tot[b] / cnt[b]:.3f}  share of excess {(tot[b] - cnt[b]) / max(1, sum(tot.values()) - sum(cnt.values())):.3f}")


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "analyse":
    analyse()

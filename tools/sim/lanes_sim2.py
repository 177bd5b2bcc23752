#!/usr/bin/env python
"""CPU model of the end-to-end ("quota") lane-stream layout: bands laid end to end in padded group space,
cut into 32 equal quotas; per-step bank-group matching.  Reports steps, padding, wavefronts per access."""
import sys
import numpy as np
from lanes_sim import make_stream, max_matching, wavefronts


def build_e2e(entries, order="band", perm=None, stagger=False, scarce=False, split_pools=True):
    K = len(entries)
    n = [len(e) for e in entries]
    bands = list(range(K))
    if order == "size":
        bands.sort(key=lambda k: -n[k])
    groups = [(n[k] + 3) // 4 for k in bands]
    M = sum(groups)
    Q = -(-M // 32)
    # lane quotas: consecutive group ranges
    runs = [[] for _ in range(32)]       # (band, ngroups_real_entries) pieces: (band, n_entries)
    lane, room = 0, Q
    for k, g in zip(bands, groups):
        ent = n[k]
        while g > 0:
            take = min(g, room)
            ne = min(ent, 4 * take)
            runs[lane].append([k, ne, take])
            ent -= ne
            g -= take
            room -= take
            if room == 0:
                lane += 1
                room = Q
    if perm is not None:
        runs = [runs[i] for i in perm]
    pools = {k: list(entries[k]) for k in range(K)}
    if split_pools:           # every piece owns a contiguous sub-range of its band's entries
        for l in range(32):
            for r in runs[l]:
                k, ne, take = r
                r.append(pools[k][:ne])
                pools[k] = pools[k][ne:]
    cur = [None] * 32         # [band, entries left in piece, groups left, (own pool)]
    steps = 0
    wf = []
    left = sum(n)
    while steps < 4 * Q:
        if steps % 4 == 0:
            for l in range(32):
                if cur[l] is None or cur[l][2] == 0:
                    cur[l] = runs[l].pop(0) if runs[l] else None
                if cur[l] is not None:
                    cur[l][2] -= 1
        chosen = [None] * 32
        taken = set()
        for q in range(4):
            ls = list(range(8 * q, 8 * q + 8))
            cands = []
            for l in ls:
                if cur[l] is None or cur[l][1] == 0:
                    cands.append([])
                    continue
                # a piece may take any entry of the band still unplaced (pieces of one band share the pool)
                cnt = {}
                for p in (cur[l][3] if split_pools else pools[cur[l][0]]):
                    if p not in taken:
                        cnt[p & 7] = cnt.get(p & 7, 0) + 1
                cands.append(sorted(cnt, key=lambda c: -cnt[c]))
            if scarce:
                idx = sorted(range(8), key=lambda i: (sum(1 for _ in cands[i]) == 0, len(cur[ls[i]][3]) if cur[ls[i]] is not None and split_pools else 99))
                m0 = max_matching([cands[i] for i in idx])
                m = {idx[a]: c for a, c in m0.items()}
                order_l = [(ls[i], i) for i in idx]
            else:
                m = max_matching(cands)
                order_l = [(l, i) for i, l in enumerate(ls)]
            for l, i in order_l:
                if cur[l] is None or cur[l][1] == 0:
                    continue
                pool = cur[l][3] if split_pools else pools[cur[l][0]]
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
                    cur[l][1] -= 1
                    left -= 1
        wf.append(wavefronts(chosen))
        steps += 1
    return steps, wf, sum(n), left


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    for order, scarce in (("band", False), ("band", True)):
        st, wfs, ents, lefts = [], [], [], 0
        for i in range(n):
            e = make_stream(np.random.default_rng(i))
            s, wf, t, left = build_e2e(e, order=order, scarce=scarce)
            st.append(s)
            ents.append(t)
            lefts += left
            wfs.append(sum(wf) / max(1, sum(1 for w in wf if w)))
        print(f"e2e order={order} scarce={scarce}: steps {np.mean(st):.2f} (ideal {np.mean(ents) / 32:.2f}), padding "
              f"{1 - np.sum(ents) / (32 * np.sum(st)):.4f}, wavefronts/access {np.mean(wfs):.3f}, unplaced {lefts}")


if __name__ == "__main__":
    main()

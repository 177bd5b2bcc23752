#!/usr/bin/env python
"""End-to-end quota layout, step selection modelled on the CUDA builder: per-quarter bank-group matching gives
every lane a wanted class; lanes propose candidates; pixel collisions are resolved by priority, losers advance."""
import sys
import numpy as np
from lanes_sim import make_stream, max_matching as exact_matching, wavefronts

GREEDY = False


def max_matching(cands):
    """cands[i]: classes in preference order.  GREEDY: most-constrained lane first, takes its first free class."""
    if not GREEDY:
        return exact_matching(cands)
    free = set(range(8))
    res = {}
    todo = [i for i in range(len(cands)) if cands[i]]
    while todo:
        best = min(todo, key=lambda i: (sum(1 for c in cands[i] if c in free), i))
        todo.remove(best)
        for c in cands[best]:
            if c in free:
                res[best] = c
                free.discard(c)
                break
    return res



def plan_runs(n, Q=None, tail_slack=True, stretch=True, spread=False):
    """Bands laid end to end in padded group space and cut into 32 quotas.  tail_slack: a band's entries are
    assigned to its pieces from the back (the piece at the end of a quota holds the partial group), and quotas
    are evened out (Q or Q-1 groups; a short lane's last piece is stretched by one group)."""
    K = len(n)
    groups = [(x + 3) // 4 for x in n]
    M = sum(groups)
    if Q is None:
        Q = -(-M // 32)
    short = 32 * Q - M if tail_slack else 0          # lanes that get Q-1 groups of their own
    if spread:
        quota = [Q - 1 if (tail_slack and ((l + 1) * short) // 32 != (l * short) // 32) else Q for l in range(32)]
    else:
        quota = [Q - 1 if (tail_slack and l >= 32 - short) else Q for l in range(32)]
    if not tail_slack:
        quota = [Q] * 32
    runs = [[] for _ in range(32)]
    lane, room = 0, quota[0]
    for k, g in enumerate(groups):
        ent = n[k]
        first = True
        pieces = []
        while g > 0:
            while room == 0:
                lane += 1
                room = quota[lane]
            take = min(g, room)
            pieces.append([lane, take])
            g -= take
            room -= take
        # entries: from the back, later pieces take full groups
        if tail_slack:
            rem = ent
            cnts = [0] * len(pieces)
            for i in range(len(pieces) - 1, 0, -1):
                cnts[i] = min(rem, 4 * pieces[i][1])
                rem -= cnts[i]
            cnts[0] = rem
        else:
            rem = ent
            cnts = []
            for ln, take in pieces:
                c = min(rem, 4 * take)
                cnts.append(c)
                rem -= c
        for i, (ln, take) in enumerate(pieces):
            runs[ln].append(dict(band=k, ne=cnts[i], ng=take, cont=i > 0))
    if tail_slack:
        for l in range(32):
            if stretch and quota[l] < Q and runs[l]:
                runs[l][-1]["ng"] += 1
    return runs, Q


def build(entries, priority="lane", shared=True, qbal=False, use_slack=False, stretch=True, spread=False):
    K = len(entries)
    n = [len(e) for e in entries]
    runs, Q = plan_runs(n, stretch=stretch, spread=spread)
    pool = {k: list(entries[k]) for k in range(K)}      # unplaced entries per band (shared by its pieces)
    cur = [None] * 32
    wf = []
    unplaced = 0
    out = [[None] * (4 * Q) for _ in range(32)]
    fails = []
    for step in range(4 * Q):
        if step % 4 == 0:
            for l in range(32):
                if cur[l] is None or cur[l]["ng"] == 0:
                    if cur[l] is not None and cur[l]["ne"] > 0:
                        unplaced += cur[l]["ne"]
                        fails.append((l, cur[l]["band"], cur[l]["t0"], step, cur[l]["ne"]))
                    cur[l] = runs[l].pop(0) if runs[l] else None
                    if cur[l] is not None:
                        cur[l]["t0"] = step
                if cur[l] is not None:
                    cur[l]["ng"] -= 1
                    cur[l]["slots"] = 4 * (cur[l]["ng"] + 1)
        active = [l for l in range(32) if cur[l] is not None and cur[l]["ne"] > 0]
        want = {}
        for q in range(4):
            ls = [l for l in active if l // 8 == q]
            cands = []
            qtot = {}
            for l in ls:
                for p in pool[cur[l]["band"]]:
                    qtot[p & 7] = qtot.get(p & 7, 0) + 1
            for l in ls:
                cnt = {}
                for p in pool[cur[l]["band"]]:
                    cnt[p & 7] = cnt.get(p & 7, 0) + 1
                if qbal:
                    cands.append(sorted(cnt, key=lambda c: (-qtot[c], -cnt[c])))
                else:
                    cands.append(sorted(cnt, key=lambda c: -cnt[c]))
            if use_slack:
                idx = sorted(range(len(ls)), key=lambda i: (cur[ls[i]]["slots"] - cur[ls[i]]["ne"], len(pool[cur[ls[i]]["band"]])))
                m0 = max_matching([cands[i] for i in idx])
                m = {idx[a]: c for a, c in m0.items()}
            else:
                m = max_matching(cands)
            for i, l in enumerate(ls):
                want[l] = m.get(i, -1)
        # candidate lists: wanted class first, then the rest
        cl = {}
        for l in active:
            pl = pool[cur[l]["band"]]
            a = [p for p in pl if (p & 7) == want[l]]
            b = [p for p in pl if (p & 7) != want[l]]
            if use_slack and cur[l]["slots"] - cur[l]["ne"] > 0:
                b = []          # a lane with slack skips rather than take a class the matching did not give it
            cl[l] = a + b
        pos = {l: 0 for l in active}
        won = {}
        und = set(active)
        taken = set()
        while und:
            prop = {}
            for l in list(und):
                while pos[l] < len(cl[l]) and cl[l][pos[l]] in taken:
                    pos[l] += 1
                if pos[l] >= len(cl[l]):
                    und.discard(l)
                    continue
                prop.setdefault(cl[l][pos[l]], []).append(l)
            for p, ls in prop.items():
                if priority == "lane":
                    w = min(ls)
                else:
                    w = min(ls, key=lambda l: (cur[l]["slots"] - cur[l]["ne"], len(cl[l]) - pos[l], l))
                won[w] = p
                taken.add(p)
                und.discard(w)
        chosen = [None] * 32
        for l, p in won.items():
            chosen[l] = p
            pool[cur[l]["band"]].remove(p)
            cur[l]["ne"] -= 1
            out[l][step] = (cur[l]["band"], p)
        for l in range(32):
            if cur[l] is not None:
                cur[l]["slots"] -= 1
        wf.append(wavefronts(chosen))
    for l in range(32):
        if cur[l] is not None and cur[l]["ne"] > 0:
            unplaced += cur[l]["ne"]
            fails.append((l, cur[l]["band"], cur[l]["t0"], 4 * Q, cur[l]["ne"]))
        for r in runs[l]:
            unplaced += r["ne"]
    # repair: unplaced entries go to padding slots of their piece (directly or by a swap inside the piece)
    repaired = 0
    T = 4 * Q
    used = [set(out[l][t][1] for l in range(32) if out[l][t] is not None) for t in range(T)]
    for (l, band, t0, t1, ne) in fails:
        for _ in range(ne):
            p = pool[band][0]
            done = False
            pads = [t for t in range(t0, t1) if out[l][t] is None]
            for t in pads:
                if p not in used[t]:
                    out[l][t] = (band, p); used[t].add(p); done = True
                    break
            if not done:
                for t in range(t0, t1):
                    if out[l][t] is None or p in used[t]:
                        continue
                    p2 = out[l][t][1]
                    for t2 in pads:
                        if p2 not in used[t2]:
                            out[l][t2] = (band, p2); used[t2].add(p2)
                            used[t].discard(p2); out[l][t] = (band, p); used[t].add(p)
                            done = True
                            break
                    if done:
                        break
            if done:
                pool[band].pop(0)
                repaired += 1
    unplaced -= repaired
    wf = [wavefronts([out[l][t][1] if out[l][t] is not None else None for l in range(32)]) for t in range(T)]
    return 4 * Q, wf, sum(n), unplaced


def main():
    global GREEDY
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    GREEDY = len(sys.argv) > 2 and sys.argv[2] == "greedy"
    for prio, qbal, us, stretch, spread in (("slack", False, True, True, False), ("slack", False, True, True, True), ("slack", False, True, False, True), ("slack", False, True, False, False)):
        st, wfs, ents, un = [], [], [], 0
        for i in range(n):
            e = make_stream(np.random.default_rng(i))
            s, wf, t, u = build(e, priority=prio, qbal=qbal, use_slack=us, stretch=stretch, spread=spread)
            st.append(s); ents.append(t); un += u
            wfs.append(sum(wf) / max(1, sum(1 for w in wf if w)))
        print(f"e2e stretch={stretch} spread={spread} use_slack={us}: steps {np.mean(st):.2f} (ideal {np.mean(ents) / 32:.2f}) padding {1 - np.sum(ents) / (32 * np.sum(st)):.4f} "
              f"wavefronts/access {np.mean(wfs):.3f} unplaced {un} ({un / n:.2f}/stream)")


if __name__ == "__main__":
    main()

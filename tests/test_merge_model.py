"""The batched merge of DESIGN.md §4 == the reference's sequential inserts.

The CUDA kernel evaluates all neighbours of an expanded node and then merges them into the sorted buffer
in one step.  This test restates that merge in plain Python (same steps, same arithmetic on keys) and
checks, on random batches full of exact ties and repeated ids, that buffer contents, visited flags and
the best-unvisited cursor equal what the oracle's `SearchBuffer` restatement produces by inserting the
same candidates one by one in adjacency order (the oracle itself is pinned to the reference's
known-answer sequence in test_search_buffer.py)."""
import numpy as np
import pytest

from test_search_buffer import Buf


def batched_merge(keys, ids, visited, cursor, cand_keys, cand_ids, capacity):
    """keys/ids/visited: current buffer (sorted by key, '<' order; IP/cosine negate distances).
    Returns the new (keys, ids, visited, cursor)."""
    size = len(keys)
    full = size == capacity
    back = keys[-1] if size else 0.0
    surv = []   # (key, id, insertion point), in adjacency order
    for d, i in zip(cand_keys, cand_ids):
        if full and back < d:                        # can_skip against the back at the start of the hop
            continue
        ipos = sum(1 for k in keys if not (d < k))   # lower_bound with !cmp(d, other)
        j, dup = ipos, False
        while j > 0:                                 # duplicate-id scan over the equal-key run
            j -= 1
            if keys[j] < d:
                break
            if ids[j] == i:
                dup = True
                break
        if not dup:
            surv.append((d, i, ipos))
    if not surv:
        return keys, ids, visited, cursor
    S = len(surv)
    # final position = insertion point + stable rank among survivors (key, then adjacency order)
    fps = []
    for si, (md, _, ipos) in enumerate(surv):
        rank = sum(1 for s, (ds, _, _) in enumerate(surv) if ds < md or (not (md < ds) and s < si))
        fps.append(ipos + rank)
    out_k, out_i, out_v = [None] * (size + S), [None] * (size + S), [None] * (size + S)
    for j in range(size):                            # old entries shift by the survivors at or before them
        shift = sum(1 for (_, _, ipos) in surv if ipos <= j)
        out_k[j + shift], out_i[j + shift], out_v[j + shift] = keys[j], ids[j], visited[j]
    for (d, i, _), fp in zip(surv, fps):
        assert out_k[fp] is None
        out_k[fp], out_i[fp], out_v[fp] = d, i, False
    new_size = min(size + S, capacity)
    minpos = min(ipos for _, _, ipos in surv)
    return out_k[:new_size], out_i[:new_size], out_v[:new_size], min(cursor, minpos)


@pytest.mark.parametrize("greater", [False, True])
@pytest.mark.parametrize("window,capacity", [(1, 1), (5, 5), (4, 9), (32, 32)])
def test_batched_merge_equals_sequential_inserts(oracle, greater, window, capacity):
    rng = np.random.default_rng(capacity * 7 + window + greater)
    sign = -1.0 if greater else 1.0
    for trial in range(60):
        seq = Buf(oracle, window, capacity, greater)
        dist_of = {}

        def dist(i):   # same id => same distance, few distinct values => many ties
            return dist_of.setdefault(int(i), float(rng.integers(0, 9)))
        ep = int(rng.integers(0, 60))
        seq.push_back(ep, dist(ep))
        keys, ids, visited, cursor = [sign * dist(ep)], [ep], [False], 0
        for hop in range(25):
            upper = min(len(keys), window)
            while cursor < upper and visited[cursor]:
                cursor += 1
            done = cursor >= upper
            assert bool(seq.l.oracle_buffer_done(seq.h)) == done
            if done:
                break
            node = seq.l.oracle_buffer_next(seq.h)
            assert node == ids[cursor]
            visited[cursor] = True
            cursor += 1
            nbrs = rng.integers(0, 60, size=int(rng.integers(0, 12)))
            nbrs = list(dict.fromkeys(int(x) for x in nbrs))          # adjacency rows hold distinct ids
            for i in nbrs:                                             # reference: one insert per neighbour
                seq.insert(i, dist(i))
            keys, ids, visited, cursor = batched_merge(keys, ids, visited, cursor, [sign * dist(i) for i in nbrs],
                                                       nbrs, capacity)
            got = seq.entries()
            assert [(i, sign * k, v) for k, i, v in zip(keys, ids, visited)] == got, (trial, hop)


def gather_merge(keys, ids, visited, cursor, cand_keys, cand_ids, capacity, lanes):
    """The lean kernel's merge (csrc/search_fast.cuh), restated step by step with a warp of `lanes` lanes:
    candidates go in groups of `lanes`; per group can_skip against the back *at the start of the group*,
    insertion point by binary search, duplicate scan, stable rank among the group's survivors, then a GATHER
    over final slots: the survivors' final positions fp form one bit mask per slot of `lanes` entries; final
    slot f takes the survivor with fp == f or the old entry f - #{fp < f}; blocks of 4 slots are rewritten
    in place, top-down, each block read completely before it is written."""
    keys, ids, visited = list(keys), list(ids), list(visited)
    for r0 in range(0, len(cand_keys), lanes):
        ck, ci = cand_keys[r0:r0 + lanes], cand_ids[r0:r0 + lanes]
        size = len(keys)
        full = size == capacity
        back = keys[-1] if full else 0.0
        surv = [not (full and back < d) for d in ck]
        if not any(surv):
            continue
        ipos = []
        for d in ck:                                   # every lane searches (harmless for non-survivors)
            pos, step = 0, (1 << (size.bit_length() - 1)) if size else 0
            while step:
                j = pos + step
                if j <= size and not (d < keys[j - 1]):
                    pos = j
                step >>= 1
            ipos.append(pos)
        for l, (d, i) in enumerate(zip(ck, ci)):
            if surv[l]:
                j = ipos[l]
                while j > 0:
                    j -= 1
                    if keys[j] < d:
                        break
                    if ids[j] == i:
                        surv[l] = False
                        break
        lanes_s = [l for l in range(len(ck)) if surv[l]]
        if not lanes_s:
            continue
        S = len(lanes_s)
        skc = [ck[l] for l in lanes_s]                 # compacted keys, adjacency order
        fp, sst = {}, [None] * S
        for t_me, l in enumerate(lanes_s):
            d = ck[l]
            rank = sum(1 for t2, ds in enumerate(skc) if ds < d or (ds == d and t2 < t_me))
            fp[l] = ipos[l] + rank
            assert sst[rank] is None
            sst[rank] = (d, ci[l])
        minpos = min(ipos[l] for l in lanes_s)
        newsize = min(size + S, capacity)
        top, lo = (newsize - 1) // lanes, minpos // lanes
        top4, lo4 = top // 4, lo // 4
        above = sum(1 for l in lanes_s if fp[l] >= 4 * lanes * (top4 + 1))
        buf = [(k, i, v) for k, i, v in zip(keys, ids, visited)] + [None] * (capacity + 4 * lanes)
        for b4 in range(top4, lo4 - 1, -1):
            writes = []
            for u in range(3, -1, -1):
                sl = 4 * b4 + u
                Ms = [any(fp[l] == sl * lanes + b for l in lanes_s) for b in range(lanes)]
                if lo <= sl <= top:
                    for lane in range(lanes):
                        f = sl * lanes + lane
                        before = S - (above + sum(Ms[lane:]))
                        if f < newsize and f >= minpos:
                            if Ms[lane]:
                                writes.append((f, (sst[before][0], sst[before][1], False)))
                            else:
                                assert 0 <= f - before < size
                                writes.append((f, buf[f - before]))
                above += sum(Ms)
            for f, e in writes:                        # after the block's reads (the __syncwarp)
                buf[f] = e
        keys = [e[0] for e in buf[:newsize]]
        ids = [e[1] for e in buf[:newsize]]
        visited = [e[2] for e in buf[:newsize]]
        cursor = min(cursor, minpos)
    return keys, ids, visited, cursor


@pytest.mark.parametrize("greater", [False, True])
@pytest.mark.parametrize("window,capacity,lanes", [(1, 1, 4), (5, 5, 2), (4, 9, 4), (32, 32, 4), (20, 37, 3), (40, 40, 32),
                                                   (130, 130, 32), (7, 300, 32)])
def test_gather_merge_equals_sequential_inserts(oracle, greater, window, capacity, lanes):
    rng = np.random.default_rng(capacity * 11 + window + greater + lanes)
    sign = -1.0 if greater else 1.0
    big = capacity > 100
    for trial in range(8 if big else 40):
        seq = Buf(oracle, window, capacity, greater)
        dist_of = {}
        nids = 600 if big else 60

        def dist(i):   # same id => same distance, few distinct values => many ties
            return dist_of.setdefault(int(i), float(rng.integers(0, 40 if big else 9)))
        ep = int(rng.integers(0, nids))
        seq.push_back(ep, dist(ep))
        keys, ids, visited, cursor = [sign * dist(ep)], [ep], [False], 0
        for hop in range(40 if big else 25):
            upper = min(len(keys), window)
            while cursor < upper and visited[cursor]:
                cursor += 1
            done = cursor >= upper
            assert bool(seq.l.oracle_buffer_done(seq.h)) == done
            if done:
                break
            node = seq.l.oracle_buffer_next(seq.h)
            assert node == ids[cursor]
            visited[cursor] = True
            cursor += 1
            nbrs = rng.integers(0, nids, size=int(rng.integers(0, 70 if big else 12)))
            nbrs = list(dict.fromkeys(int(x) for x in nbrs))          # adjacency rows hold distinct ids
            for i in nbrs:                                             # reference: one insert per neighbour
                seq.insert(i, dist(i))
            keys, ids, visited, cursor = gather_merge(keys, ids, visited, cursor, [sign * dist(i) for i in nbrs],
                                                      nbrs, capacity, lanes)
            got = seq.entries()
            assert [(i, sign * k, v) for k, i, v in zip(keys, ids, visited)] == got, (trial, hop)

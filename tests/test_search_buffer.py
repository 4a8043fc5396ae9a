"""SearchBuffer semantics of the oracle, pinned by the reference's own unit test.

Replays the known-answer insert sequence of /root/reference/tests/svs/index/vamana/search_buffer.cpp:382-519
(begin / middle / end x full / non-full x duplicate id, return values included), the sort section (:521-541), and
fuzzes against an independent list model in the spirit of `SearchBufferReference` (:74-244) for std::less and
std::greater."""
import ctypes as C

import numpy as np
import pytest


class Buf:
    def __init__(self, lib, window, capacity, greater=False):
        self.l = lib.lib
        self.l.oracle_buffer_new.restype = C.c_void_p
        self.l.oracle_buffer_new.argtypes = [C.c_size_t, C.c_size_t, C.c_int]
        for name in ("insert", "size", "best_unvisited"):
            getattr(self.l, f"oracle_buffer_{name}").restype = C.c_size_t
        self.l.oracle_buffer_insert.argtypes = [C.c_void_p, C.c_uint32, C.c_float]
        self.l.oracle_buffer_push_back.argtypes = [C.c_void_p, C.c_uint32, C.c_float]
        self.l.oracle_buffer_get.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        self.l.oracle_buffer_set_visited.argtypes = [C.c_void_p, C.c_size_t]
        for name in ("size", "best_unvisited", "done", "next", "clear", "free", "sort"):
            getattr(self.l, f"oracle_buffer_{name}").argtypes = [C.c_void_p]
        self.l.oracle_buffer_next.restype = C.c_uint32
        self.h = self.l.oracle_buffer_new(window, capacity, int(greater))

    def __del__(self):
        self.l.oracle_buffer_free(self.h)

    def insert(self, id, d):
        return self.l.oracle_buffer_insert(self.h, id, d)

    def push_back(self, id, d):
        self.l.oracle_buffer_push_back(self.h, id, d)

    def __len__(self):
        return self.l.oracle_buffer_size(self.h)

    def __getitem__(self, i):
        id, d, v = C.c_uint32(), C.c_float(), C.c_int()
        self.l.oracle_buffer_get(self.h, i, C.byref(id), C.byref(d), C.byref(v))
        return (id.value, d.value, bool(v.value))

    def entries(self):
        return [self[i] for i in range(len(self))]


def test_known_answer_insert_sequence(oracle):
    b = Buf(oracle, 5, 5)
    b.push_back(1, 10)
    assert b[0] == (1, 10, False)
    assert b.insert(2, 20) == 1 and b.entries() == [(1, 10, False), (2, 20, False)]               # 1a
    b.l.oracle_buffer_set_visited(b.h, 1)
    assert b.insert(2, 20) == len(b) + 1 and b.entries() == [(1, 10, False), (2, 20, True)]        # 1b
    assert b.insert(3, 5) == 0 and b.entries() == [(3, 5, False), (1, 10, False), (2, 20, True)]  # 2a
    assert b.insert(3, 5) == len(b) + 1 and len(b) == 3                                            # 2b
    assert b.insert(4, 15) == 2                                                                    # 3a
    assert b.entries() == [(3, 5, False), (1, 10, False), (4, 15, False), (2, 20, True)]
    assert b.insert(4, 15) == len(b) + 1 and len(b) == 4                                           # 3b
    assert b.insert(5, 30) == 4 and len(b) == 5                                                    # prep for 4
    assert b.insert(6, 1000) == 5 and len(b) == 5                                                  # 4: skipped -> size()
    assert b.entries() == [(3, 5, False), (1, 10, False), (4, 15, False), (2, 20, True), (5, 30, False)]
    assert b.insert(7, 1) == 0                                                                     # 5a
    assert b.entries() == [(7, 1, False), (3, 5, False), (1, 10, False), (4, 15, False), (2, 20, True)]
    assert b.insert(7, 1) == len(b) + 1                                                            # 5b
    assert b.insert(8, 8) == 2                                                                     # 6a
    assert b.entries() == [(7, 1, False), (3, 5, False), (8, 8, False), (1, 10, False), (4, 15, False)]
    assert b.insert(8, 8) == len(b) + 1                                                            # 6b
    b.l.oracle_buffer_clear(b.h)
    assert len(b) == 0


@pytest.mark.parametrize("greater", [False, True])
def test_sort_section(oracle, greater):
    b = Buf(oracle, 5, 5, greater)
    for id, d in ((1, 100), (2, 10), (3, 50)):
        b.push_back(id, d)
    b.l.oracle_buffer_sort(b.h)
    want = [(1, 100), (3, 50), (2, 10)] if greater else [(2, 10), (3, 50), (1, 100)]
    assert [(i, d) for i, d, _ in b.entries()] == want


@pytest.mark.parametrize("greater", [False, True])
@pytest.mark.parametrize("window,capacity", [(4, 4), (3, 7), (16, 16)])
def test_fuzz_against_list_model(oracle, greater, window, capacity):
    """Random inserts (few distinct distances => many ties and repeated ids) interleaved with next():
    contents, visited flags, best_unvisited and done() must equal a plain-Python model."""
    rng = np.random.default_rng(window * 31 + capacity + greater)
    better = (lambda x, y: x > y) if greater else (lambda x, y: x < y)
    b = Buf(oracle, window, capacity, greater)
    model, best = [], 0        # model entries: [id, dist, visited]
    dist_of = {}
    for step in range(3000):
        if rng.random() < 0.25 and not b.l.oracle_buffer_done(b.h):
            upper = min(len(model), window)
            node = model[best]
            node[2] = True
            best += 1
            while best != upper and model[best][2]:
                best += 1
            assert b.l.oracle_buffer_next(b.h) == node[0]
        else:
            id = int(rng.integers(0, 40))
            d = dist_of.setdefault(id, float(rng.integers(0, 12)))   # same id => same distance
            full = len(model) == capacity
            if full and better(model[-1][1], d):
                want = len(model)
            else:
                pos = 0
                while pos < len(model) and not better(d, model[pos][1]):
                    pos += 1
                j, dup = pos, False
                while j > 0:
                    j -= 1
                    if better(model[j][1], d):
                        break
                    if model[j][0] == id:
                        dup = True
                        break
                if dup:
                    want = len(model) + 1
                else:
                    model.insert(pos, [id, d, False])
                    del model[capacity:]
                    best = min(best, pos)
                    want = pos
            assert b.insert(id, d) == want, step
        assert b.entries() == [tuple(e) for e in model], step
        assert b.l.oracle_buffer_best_unvisited(b.h) == best
        assert bool(b.l.oracle_buffer_done(b.h)) == (best == min(len(model), window))

"""Executable statements of the two rules the level-synchronous walk relies on for large frontiers ("K5, large frontiers";
kernels.cuh bfs_search_pairs and the fallback rounds of lineage_bfs_kernel).  Pure numpy / Python: they check the ARGUMENT, the CUDA code
is checked against the oracle and against its own dense variant on the GPU (tests/test_gpu_parity.py)."""
import numpy as np


def _last_within_tolerance(v, eps=1e-12):
    """max_by(partial_cmp) of lineage.rs:156-164 with the device's tie rule: the LAST child within eps relative of the maximum."""
    best = v.max()
    thr = best - abs(best) * eps
    return int(np.nonzero(v >= thr)[0][-1])


def _one_pass(v, order, eps=1e-12):
    """The one-pass form: children arrive in any order (warps run concurrently); a child within the tolerance of the RUNNING maximum is a
    record; the record with the largest index is the answer if it passes the final test, else the chain is scanned again."""
    run = -np.inf
    last_rec = -1
    for i in order:
        run = max(run, v[i])
        if v[i] >= run - abs(run) * eps:
            last_rec = max(last_rec, int(i))
    best = v.max()
    thr = best - abs(best) * eps
    if v[last_rec] >= thr:
        return last_rec, False
    return _last_within_tolerance(v, eps), True  # second pass


def test_one_pass_argmax_equals_two_pass():
    rng = np.random.default_rng(5)
    slow = 0
    for case in range(3000):
        n = int(rng.integers(1, 200))
        kind = case % 4
        if kind == 0:
            v = rng.random(n)
        elif kind == 1:  # many exact ties (identical hit counts)
            v = rng.choice([0.0, 1e-9, 3e-7, 0.25], n)
        elif kind == 2:  # ties up to rounding noise
            v = 0.125 * (1.0 + rng.integers(-3, 4, n) * 1.1e-16)
        else:  # values in the band right below the tolerance: the case that needs the second pass
            v = 0.5 * (1.0 - rng.integers(0, 4, n) * 0.7e-12)
        order = rng.permutation(n)
        got, second = _one_pass(v, order)
        assert got == _last_within_tolerance(v)
        slow += second
    assert slow > 0  # the band case is exercised


def _significant_dense(prefix, bounds):
    conf = prefix[bounds[1:]] - prefix[bounds[:-1]]
    return set(np.nonzero(np.round(conf * 100.0) != 0)[0].tolist())


def _significant_by_search(prefix, bounds, leaf=4, floor=0.004999):
    """Mass-pruned search: a run of children is cut into (up to) 32 sub-runs; a sub-run whose mass stays below the floor is dropped,
    runs of at most `leaf` children are evaluated child by child."""
    out, probes = set(), 0
    todo = [(0, len(bounds) - 1)]
    while todo:
        a, b = todo.pop()
        if b - a <= leaf:
            conf = prefix[bounds[a + 1: b + 1]] - prefix[bounds[a: b]]
            out |= {a + int(i) for i in np.nonzero(np.round(conf * 100.0) != 0)[0]}
            continue
        step = (b - a + 31) // 32
        for sa in range(a, b, step):
            sb = min(sa + step, b)
            probes += 1
            if prefix[bounds[sb]] - prefix[bounds[sa]] >= floor:
                todo.append((sa, sb))
    return out, probes


def test_mass_pruned_search_finds_every_significant_child():
    rng = np.random.default_rng(6)
    for case in range(300):
        n_refs = int(rng.integers(50, 20000))
        # probability profile: flat, peaked, or a few plateaus
        p = rng.random(n_refs) ** (1 if case % 3 == 0 else 40 if case % 3 == 1 else 8)
        if case % 5 == 0:
            p[rng.integers(0, n_refs, 3)] += p.sum()  # near one-hot
        p /= p.sum()
        prefix = np.concatenate([[0.0], np.cumsum(p)])
        n_children = int(rng.integers(1, min(n_refs, 5000) + 1))
        cuts = np.sort(rng.choice(np.arange(1, n_refs), n_children - 1, replace=False)) if n_children > 1 else np.zeros(0, np.int64)
        bounds = np.concatenate([[0], cuts, [n_refs]]).astype(np.int64)
        want = _significant_dense(prefix, bounds)
        got, probes = _significant_by_search(prefix, bounds)
        assert got == want
        assert len(want) <= 200 + 1  # children of at least 0.005 each
        if n_children >= 2000:
            assert probes < n_children  # fewer probes than children on a large frontier

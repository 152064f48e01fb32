"""Host-side logic of the multi-GPU paths on CPU: query partitioning, shard cuts, and the per-query merge of
reference-sharded result lines -- the latter also across two processes over the gloo backend (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

from raxtax_b200 import capi, dist as rdist, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_queries_covers_everything():
    for n, w in [(10, 1), (10, 3), (7, 8), (200000, 8), (0, 4)]:
        spans = [rdist.partition_queries(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_shard_cuts():
    cuts = rdist.shard_cuts(8_000_000, 8)
    assert cuts[0] == 0 and cuts[-1] == 8_000_000 and len(cuts) == 9
    assert np.all(np.diff(cuts.astype(np.int64)) > 0) and np.all(cuts[1:-1] % 32 == 0)
    assert list(rdist.shard_cuts(5, 3)) == [0, 1, 2, 5]


def _oracle_case(oracle):
    ds = synth.generate("tiny", measure=False)
    ot = oracle.Tree.new(ds.ref_lineages, [ds.ref_seq(i) for i in range(ds.n_refs)])
    o = ot.classify(ds.query_off, ds.query_codes, skip_exact=False, raw_conf=True, threads=2, chunk_size=16)
    o_ovr = ot.classify(ds.query_off, ds.query_codes, skip_exact=False, raw_conf=False, threads=2, chunk_size=16)
    return ds, ot, o, o_ovr


def _to_output(res, nq, K, ML, keep):
    """ClassifyOutput holding the subset `keep` (boolean per oracle line) of an oracle Results object."""
    begin, first, nlev, conf, local = [0], [], [], [], []
    glob = np.zeros(nq)
    for q in range(nq):
        for i in np.nonzero(res.query == q)[0]:
            glob[q] = res.glob[i]
            if keep[i]:
                first.append(res.first_ref[i]); nlev.append(res.nlev[i]); conf.append(res.conf[i, :ML]); local.append(res.local[i])
        begin.append(len(first))
    return capi.ClassifyOutput(K.copy(), np.asarray(begin, np.uint32), glob, np.asarray(first, np.uint32), np.asarray(nlev, np.uint8),
                               np.asarray(conf, np.float64).reshape(len(first), ML), np.asarray(local, np.float64))


def _assert_same(merged, res, nq):
    for q in range(nq):
        a = res.for_query(q)
        b = merged.for_query(q)
        assert len(a) == len(b)
        for (fa, ca, la, ga), (fb, cb, lb, gb) in zip(a, b):
            assert fa == fb and np.array_equal(ca, cb) and la == lb and ga == gb


def test_merge_restores_reference_order(oracle):
    ds, ot, o, o_ovr = _oracle_case(oracle)
    res, nq, ML = o["results"], ds.n_queries, 6
    cuts = rdist.shard_cuts(ot.num_tips, 3)
    owner = np.searchsorted(cuts, res.first_ref, side="right") - 1
    outs = [_to_output(res, nq, o["K"], ML, owner == r) for r in range(3)]
    ht = capi.Tree.new(ds.ref_lineages, ds.ref_off, ds.ref_codes)
    eo, eids = ht.exact_batch(ds.query_off, ds.query_codes)
    ref_levels = ht.index_arrays()["ref_levels"]
    merged = rdist.merge_shard_results(outs[::-1], eo, eids, ref_levels, raw_conf=True)  # rank order must not matter
    _assert_same(merged, res, nq)
    merged = rdist.merge_shard_results(outs, eo, eids, ref_levels, raw_conf=False)  # override applied after the merge
    _assert_same(merged, o_ovr["results"], nq)


def _gloo_worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist

        from oracle import oracle as orc

        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        ds, ot, o, _ = _oracle_case(orc)
        res, nq, ML = o["results"], ds.n_queries, 6
        cuts = rdist.shard_cuts(ot.num_tips, world)
        owner = np.searchsorted(cuts, res.first_ref, side="right") - 1
        mine = _to_output(res, nq, o["K"], ML, owner == rank)  # what this rank's phase 3 would emit
        lo, hi = rdist.partition_queries(nq, world, rank)
        assert hi - lo >= nq // world
        outs = rdist.gather_outputs(mine)
        ht = capi.Tree.new(ds.ref_lineages, ds.ref_off, ds.ref_codes)
        eo, eids = ht.exact_batch(ds.query_off, ds.query_codes)
        merged = rdist.merge_shard_results(outs, eo, eids, ht.index_arrays()["ref_levels"], raw_conf=True)
        _assert_same(merged, res, nq)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, "".join(traceback.format_exception(e))))


def test_sharded_merge_over_gloo_world_size_2():
    import multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(got) == [(0, "ok"), (1, "ok")], got


@pytest.mark.parametrize("seed", range(12))
def test_native_merge_equals_python_statement(seed):
    """rxh_merge_shard_results against the pure-Python statement of the same rule on random per-rank outputs: ties between confidence
    vectors, vectors of different length with an equal prefix, queries with and without a single exact match."""
    rng = np.random.default_rng(seed)
    n_ranks, nq, ML, n_refs = int(rng.integers(1, 6)), int(rng.integers(1, 40)), int(rng.integers(1, 7)), 500
    ref_levels = rng.integers(1, ML + 1, n_refs).astype(np.uint8)
    outs = []
    for _ in range(n_ranks):
        begin, first, nlev, conf, local = [0], [], [], [], []
        for q in range(nq):
            for _ in range(int(rng.integers(0, 5)) + (1 if len(outs) == 0 else 0)):  # rank 0 always contributes a line: no empty query
                n = int(rng.integers(1, ML + 1))
                c = np.zeros(ML)
                c[:n] = np.sort(rng.integers(1, 5, n))[::-1] / 100.0 + (0.9 if rng.random() < 0.3 else 0.0)  # few distinct values: many ties
                first.append(int(rng.integers(0, n_refs))); nlev.append(n); conf.append(c); local.append(float(rng.random()))
            begin.append(len(first))
        outs.append(capi.ClassifyOutput(np.zeros(nq, np.uint16), np.asarray(begin, np.uint32), rng.random(nq), np.asarray(first, np.uint32),
                                        np.asarray(nlev, np.uint8), np.asarray(conf, np.float64).reshape(len(first), ML), np.asarray(local, np.float64)))
    ne = rng.integers(0, 3, nq)
    eo = np.concatenate([[0], np.cumsum(ne)]).astype(np.uint32)
    eids = rng.integers(0, n_refs, int(eo[-1])).astype(np.uint32)
    for skip, raw in [(False, False), (True, False), (False, True)]:
        a = rdist.merge_shard_results(outs, eo, eids, ref_levels, skip_exact=skip, raw_conf=raw)
        b = rdist.merge_shard_results_py(outs, eo, eids, ref_levels, skip_exact=skip, raw_conf=raw)
        assert np.array_equal(a.result_begin, b.result_begin) and np.array_equal(a.first_ref, b.first_ref)
        assert np.array_equal(a.n_levels, b.n_levels) and np.array_equal(a.confidence, b.confidence) and np.array_equal(a.local_signal, b.local_signal)

"""GPU tests of the pipelined host driver (rxh_raxtax / rxh_raxtax_multi) and of the two batch slots of a context
(rtx_batch_slot): whatever the chunking, slot interleaving and formatting threads do, the strings handed to the sender are the ones
the one-chunk run produces."""
import os

import numpy as np
import pytest

from raxtax_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def small(ctx):
    ds = synth.generate("small", measure=False)
    ht = capi.Tree.new(ds.ref_lineages, ds.ref_off, ds.ref_codes)
    ctx.upload_tree(ht)
    return ds, ht


def _same(a, b):
    n = int(a.result_begin[-1])
    assert np.array_equal(a.result_begin, b.result_begin) and np.array_equal(a.n_kmers, b.n_kmers)
    assert np.array_equal(a.global_signal, b.global_signal)
    assert np.array_equal(a.first_ref[:n], b.first_ref[:n]) and np.array_equal(a.n_levels[:n], b.n_levels[:n])
    assert np.array_equal(a.confidence[:n], b.confidence[:n]) and np.array_equal(a.local_signal[:n], b.local_signal[:n])


def test_batch_slots_interleaved_equal_sequential(ctx, small):
    """upload A | run A | upload B | run B | download A | upload A' | run A' | download B | download A': every download returns what a
    plain rtx_classify_batch of the same queries returns (the slots share the stream, the scratch and the result pool logic)."""
    ds, ht = small
    ctx.upload_tree(ht)
    cut = [0, 100, 190, ds.n_queries]
    parts = []
    for i in range(3):
        off = (ds.query_off[cut[i]: cut[i + 1] + 1] - ds.query_off[cut[i]]).astype(np.uint64)
        codes = ds.query_codes[int(ds.query_off[cut[i]]): int(ds.query_off[cut[i + 1]])]
        eo, eids = ht.exact_batch(off, codes)
        parts.append((off, codes, eo, eids))
    ctx.batch_slot(0)
    want = [ctx.classify(*p) for p in parts]
    for sub in (0, 23):
        ctx.set_option(capi.RTX_OPT_SUB_BATCH, sub)
        ctx.batch_slot(0)
        ctx.batch_upload(*parts[0])
        ctx.batch_run()
        ctx.batch_slot(1)
        ctx.batch_upload(*parts[1])
        ctx.batch_run()
        ctx.batch_slot(0)
        _same(want[0], ctx.batch_download())
        ctx.batch_upload(*parts[2])
        ctx.batch_run()
        ctx.batch_slot(1)
        _same(want[1], ctx.batch_download())
        ctx.batch_slot(0)
        _same(want[2], ctx.batch_download())
    ctx.set_option(capi.RTX_OPT_SUB_BATCH, 0)
    ctx.batch_slot(0)


@pytest.mark.parametrize("skip,tsv", [(False, True), (True, False)])
def test_driver_chunking_does_not_change_the_output(ctx, small, skip, tsv):
    ds, ht = small
    ctx.upload_tree(ht)
    qs = capi.Queries.new(ds.query_labels, ds.query_off, ds.query_codes)
    one, logs1, warn1 = capi.raxtax(ctx, qs, ht, skip_exact_matches=skip, chunk_size=ds.n_queries, tsv=tsv)
    assert [r[0] for r in one] == ds.query_labels
    for chunk in (1, 7, 64, 0):  # 1: every query its own batch (256 trips through the 3-job ring); 0: the driver's default
        got, logs, warn = capi.raxtax(ctx, qs, ht, skip_exact_matches=skip, chunk_size=chunk, tsv=tsv)
        assert got == one, f"chunk_size={chunk}"
        assert sorted(logs) == sorted(logs1) and warn == warn1


def test_counting_sender_sees_what_a_python_sender_sees(ctx, small):
    ds, ht = small
    ctx.upload_tree(ht)
    qs = capi.Queries.new(ds.query_labels, ds.query_off, ds.query_codes)
    sent, logs, _ = capi.raxtax(ctx, qs, ht, chunk_size=50)
    a = capi.raxtax_counted(ctx, qs, ht, chunk_size=50)
    b = capi.raxtax_counted(ctx, qs, ht, chunk_size=0)
    assert a["queries"] == len(sent) == ds.n_queries
    assert a["lines"] == sum(len(s[1].split("\n")) for s in sent)
    assert a["primary_bytes"] == sum(len(s[1].encode()) for s in sent)
    assert a["label_bytes"] == sum(len(l.encode()) for l in ds.query_labels)
    assert a["log_lines"] == len(logs)
    assert a["checksum"] == b["checksum"] and a["lines"] == b["lines"]
    other = capi.Context(0)
    try:  # two contexts: arrival order differs, the order-independent checksum does not
        other.upload_tree(ht)
        c = capi.raxtax_counted([ctx, other], qs, ht, chunk_size=19)
    finally:
        other.close()
    assert c["checksum"] == a["checksum"] and c["lines"] == a["lines"] and c["queries"] == a["queries"]


def test_driver_many_queries_helpers_and_dedup(ctx, small):
    """> 512 queries switch the formatting helper threads on; the repeated queries exercise the flat de-duplication table."""
    ds, ht = small
    ctx.upload_tree(ht)
    reps = 5
    labels = [f"{l}#{r}" for r in range(reps) for l in ds.query_labels]
    offs, total = [0], 0
    lens = (ds.query_off[1:] - ds.query_off[:-1]).astype(np.int64)
    for r in range(reps):
        for n in lens:
            total += int(n)
            offs.append(total)
    codes = np.concatenate([ds.query_codes] * reps)
    qs = capi.Queries.new(labels, np.asarray(offs, np.uint64), codes)
    base, _, _ = capi.raxtax(ctx, capi.Queries.new(ds.query_labels, ds.query_off, ds.query_codes), ht, chunk_size=ds.n_queries)
    strip = lambda label, text: "\n".join(line.split("\t", 1)[1] for line in text.split("\n"))
    for chunk, env in ((0, None), (300, "1"), (300, "4")):
        if env is None:
            os.environ.pop("RXH_FORMAT_THREADS", None)
        else:
            os.environ["RXH_FORMAT_THREADS"] = env
        try:
            got, _, _ = capi.raxtax(ctx, qs, ht, chunk_size=chunk)
        finally:
            os.environ.pop("RXH_FORMAT_THREADS", None)
        assert [g[0] for g in got] == labels
        for i, g in enumerate(got):
            b = base[i % ds.n_queries]
            assert strip(g[0], g[1]) == strip(b[0], b[1]), (chunk, env, i)

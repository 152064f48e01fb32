"""Multi-GPU plumbing (torch.distributed / NCCL) around the C ABI.

Two ways to use several B200s, as BASELINE.json's north_star names them:

* query-partitioned, index replicated: no collective at all -- every rank owns a `capi.Context`, uploads the whole
  index and classifies its own slice of the queries (`partition_queries`).
* reference-sharded (databases too large to replicate): rank r holds references [cuts[r], cuts[r+1]).  Per batch:
  phase 1 (local hit counts + histograms) -> all-reduce SUM of the histograms -> phase 2 (probabilities, local
  prefixes, records of the nodes that straddle a cut) -> all-gather of the records -> phase 3 (each rank walks the
  part of the lineage tree it owns) -> the per-rank result lines are merged per query (`merge_shard_results`).

`LocalShardGroup` drives several contexts from one process (used by the single-GPU tests: several shards on one
device); `TorchShardGroup` is one context per rank with torch.distributed collectives on the library's device buffers.
"""
from __future__ import annotations

import numpy as np

from . import capi


def partition_queries(n_queries: int, world: int, rank: int):
    """Contiguous, near-equal slices (raxtax.rs:35-39: queries are independent)."""
    base, rem = divmod(n_queries, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_cuts(n_refs: int, n_shards: int) -> np.ndarray:
    """Contiguous ranges of the lineage-sorted reference ids, equal sizes rounded to 32 so that bit rows stay aligned."""
    cuts = [0]
    for r in range(1, n_shards):
        c = (n_refs * r // n_shards) // 32 * 32
        cuts.append(max(c, cuts[-1] + 1))
    cuts.append(n_refs)
    return np.asarray(cuts, np.uint64)


# ---------------------------------------------------------------------------------------------------------------
class _CudaArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def device_tensor(ptr: int, n: int, typestr: str, device: int):
    """Zero-copy torch view of a device buffer owned by the library."""
    import torch

    if n == 0 or ptr == 0:
        return torch.empty(0, device=f"cuda:{device}")
    return torch.as_tensor(_CudaArray(ptr, n, typestr), device=f"cuda:{device}")


class LocalShardGroup:
    """All shards live in this process (one context per shard, any mix of devices)."""

    def __init__(self, ctxs):
        self.ctxs = list(ctxs)

    def allreduce_hist(self):
        import torch

        ts = []
        for c in self.ctxs:
            c.synchronize()
            p, n = c.shard_hist_buffer()
            ts.append(device_tensor(p, n, "<i4", c.device))
        if not ts or ts[0].numel() == 0:
            return
        total = ts[0].clone()
        for t in ts[1:]:
            total += t.to(total.device)
        for t in ts:
            t.copy_(total.to(t.device))
        torch.cuda.synchronize()

    def allgather_records(self):
        import torch

        sends, recvs = [], []
        for c in self.ctxs:
            c.synchronize()
            (sp, sb), (rp, rb) = c.shard_records_buffers()
            sends.append(device_tensor(sp, sb, "|u1", c.device))
            recvs.append(device_tensor(rp, rb, "|u1", c.device))
        if not sends or sends[0].numel() == 0:
            return
        for r in recvs:
            r.copy_(torch.cat([s.to(r.device) for s in sends]))
        torch.cuda.synchronize()


class TorchShardGroup:
    """One context per rank; collectives over torch.distributed (NCCL on GPUs: the histograms travel over NVLink)."""

    def __init__(self, ctx, group=None):
        self.ctx = ctx
        self.group = group

    def allreduce_hist(self):
        import torch
        import torch.distributed as dist

        self.ctx.synchronize()
        p, n = self.ctx.shard_hist_buffer()
        if n == 0:
            return
        t = device_tensor(p, n, "<i4", self.ctx.device)  # counts < 2^31: int32 sums are bit-identical to uint32
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        torch.cuda.synchronize()

    def allgather_records(self):
        import torch
        import torch.distributed as dist

        self.ctx.synchronize()
        (sp, sb), (rp, rb) = self.ctx.shard_records_buffers()
        if sb == 0:
            return
        send = device_tensor(sp, sb, "|u1", self.ctx.device)
        recv = device_tensor(rp, rb, "|u1", self.ctx.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        torch.cuda.synchronize()


# ---------------------------------------------------------------------------------------------------------------
def merge_shard_results(outs, exact_off, exact_ids, ref_levels, skip_exact=False, raw_conf=False) -> capi.ClassifyOutput:
    """Merge the per-rank result lines of a reference-sharded batch (rxh_merge_shard_results in the C++ host library; the pure-Python
    statement of the same rule below, `merge_shard_results_py`, is what the tests compare it with)."""
    import ctypes as C

    L = capi.host_lib()
    nq = len(outs[0].n_kmers)
    ML = outs[0].confidence.shape[1] if outs[0].confidence.ndim == 2 else 1
    keep = []  # contiguous arrays must outlive the call

    def col(get, dtype, ctype):
        arrs = [np.ascontiguousarray(get(o), dtype) for o in outs]
        arrs = [a if a.size else np.zeros(1, dtype) for a in arrs]
        keep.append(arrs)
        return (C.POINTER(ctype) * len(outs))(*[a.ctypes.data_as(C.POINTER(ctype)) for a in arrs])

    total = int(sum(int(o.result_begin[-1]) for o in outs)) + nq
    begin, first, nlev = np.zeros(nq + 1, np.uint32), np.zeros(max(total, 1), np.uint32), np.zeros(max(total, 1), np.uint8)
    conf, local = np.zeros((max(total, 1), ML), np.float64), np.zeros(max(total, 1), np.float64)
    eo = np.ascontiguousarray(exact_off, np.uint32) if exact_off is not None else None
    ei = np.ascontiguousarray(exact_ids, np.uint32) if exact_ids is not None and len(exact_ids) else np.zeros(1, np.uint32)
    rl = np.ascontiguousarray(ref_levels, np.uint8)
    n_out = C.c_uint64(0)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t)) if a is not None else None
    rc = L.rxh_merge_shard_results(len(outs), nq, ML, col(lambda o: o.result_begin, np.uint32, C.c_uint32), col(lambda o: o.first_ref, np.uint32, C.c_uint32),
                                   col(lambda o: o.n_levels, np.uint8, C.c_uint8), col(lambda o: o.confidence, np.float64, C.c_double),
                                   col(lambda o: o.local_signal, np.float64, C.c_double), p(eo, C.c_uint32), p(ei, C.c_uint32), p(rl, C.c_uint8),
                                   int(skip_exact), int(raw_conf), p(begin, C.c_uint32), p(first, C.c_uint32), p(nlev, C.c_uint8),
                                   p(conf, C.c_double), p(local, C.c_double), total, C.byref(n_out))
    if rc != 0:
        raise capi.RtxError(capi.RTX_ERR_ASSERT, L.rxh_last_error().decode())
    n = int(n_out.value)
    return capi.ClassifyOutput(outs[0].n_kmers.copy(), begin, outs[0].global_signal.copy(), first[:n], nlev[:n], conf[:n], local[:n])


def merge_shard_results_py(outs, exact_off, exact_ids, ref_levels, skip_exact=False, raw_conf=False) -> capi.ClassifyOutput:
    """The same merge, stated in Python (test reference).

    Per query: concatenate the ranks' lines, order them like lineage.rs:93 (confidence vectors descending
    lexicographically, a longer vector first on an equal prefix; ties in depth-first order == ascending first
    reference id), then apply the one-exact-match override of raxtax.rs:73-84."""
    nq = len(outs[0].n_kmers)
    ML = outs[0].confidence.shape[1] if outs[0].confidence.ndim == 2 else 1
    first, nlev, conf, local, begin = [], [], [], [], [0]
    for q in range(nq):
        lines = []
        for o in outs:
            for i in range(int(o.result_begin[q]), int(o.result_begin[q + 1])):
                n = int(o.n_levels[i])
                key = tuple(-int(round(x * 100)) for x in o.confidence[i, :n])
                lines.append((key, int(o.first_ref[i]), n, o.confidence[i].copy(), float(o.local_signal[i])))
        if not lines:
            raise capi.RtxError(capi.RTX_ERR_ASSERT, f"query {q}: empty evaluation result (raxtax.rs:72)")

        # descending lexicographic with "the longer vector first on an equal prefix": ascending on the negated hundredths,
        # a shorter vector padded with a value larger than any negated confidence
        def sort_key(l):
            key, fr, n, _, _ = l
            return (key + (1000,) * (ML - n), fr)

        lines.sort(key=sort_key)
        ne = int(exact_off[q + 1] - exact_off[q]) if exact_off is not None else 0
        if not raw_conf and not skip_exact and ne == 1:
            idx = int(exact_ids[int(exact_off[q])])
            n = int(ref_levels[idx])
            c = np.zeros(ML)
            c[:n] = 1.0
            lines = [((), idx, n, c, lines[0][4])]
        for _, fr, n, c, l in lines:
            first.append(fr)
            nlev.append(n)
            conf.append(c)
            local.append(l)
        begin.append(len(first))
    return capi.ClassifyOutput(outs[0].n_kmers.copy(), np.asarray(begin, np.uint32), outs[0].global_signal.copy(), np.asarray(first, np.uint32),
                               np.asarray(nlev, np.uint8), np.asarray(conf, np.float64).reshape(len(first), ML), np.asarray(local, np.float64))


def classify_sharded_local(ctxs, seq_off, codes, exact_off, exact_ids, ref_levels, skip_exact=False, raw_conf=False, taps=()):
    """Reference-sharded classification with every shard in this process.  Returns (merged, per-rank outputs)."""
    group = LocalShardGroup(ctxs)
    for c in ctxs:
        c.batch_upload(seq_off, codes, exact_off, exact_ids, skip_exact=skip_exact, raw_conf=raw_conf)
        c.shard_phase(1)
    group.allreduce_hist()
    for c in ctxs:
        c.shard_phase(2)
    group.allgather_records()
    for c in ctxs:
        c.shard_phase(3)
    outs = [c.batch_download(taps=taps) for c in ctxs]
    return merge_shard_results(outs, exact_off, exact_ids, ref_levels, skip_exact, raw_conf), outs


def classify_sharded_rank(ctx, group: TorchShardGroup, seq_off, codes, exact_off, exact_ids, skip_exact=False, raw_conf=False, taps=()):
    """One rank's part of a reference-sharded batch (call on every rank); returns this rank's lines (no override)."""
    ctx.batch_upload(seq_off, codes, exact_off, exact_ids, skip_exact=skip_exact, raw_conf=raw_conf)
    ctx.shard_phase(1)
    group.allreduce_hist()
    ctx.shard_phase(2)
    group.allgather_records()
    ctx.shard_phase(3)
    return ctx.batch_download(taps=taps)


def gather_outputs(out: capi.ClassifyOutput, group=None):
    """Gather every rank's ClassifyOutput on all ranks (small: a few lines per query)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    objs = [None] * world
    dist.all_gather_object(objs, out, group=group)
    return objs

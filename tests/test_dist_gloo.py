"""world_size-2 `gloo` tests on CPU for the N>1 host logic.

What is rank-local and what crosses ranks in scm_ruv3_fit is mirrored here with NumPy stand-ins for the
device stages (the oracle's arithmetic), driven through the SAME all-reduce hook and row sharding the GPU
path uses.  The test proves the decomposition

    group sums / counts  -> all-reduce -> global column mean (stand_mean / the PCA centre)
    sums of a strided sample of <= 4096 local cells -> all-reduce -> shift a (the same vector on every rank; the Gram kernel
        reads the pre-shifted copy Y - a, scm_api.cu ruv3_fit_impl)
    local Gram of (Y - a) -> all-reduce -> minus sum_g n_g (mu_g - a)(mu_g - a)'  == Gram of the residual Y0
                                        -> minus m (ybar - a)(ybar - a)'          == centred Gram of all cells (ruvK selection)

gives every rank the single-process result, and that the hook sums raw buffers in place.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from oracle import ruv_oracle as orc
    from scmerge_b200.dist import shard_rows, torch_allreduce_hook
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hook = torch_allreduce_hook(None)

        def allreduce(arr):  # exactly what the C library does: hand the hook a raw pointer
            assert arr.flags.c_contiguous
            hook(arr.ctypes.data, arr.size, 0 if arr.dtype == np.float64 else 1, 0)

        P = orc.planted_factors(1500, 60, 24, 4, n_types=5, n_batch=3, seed=9)
        Y, g, G = P["Y"], P["types"], 5
        lo, hi = shard_rows(Y.shape[0], world, rank)
        Yl, gl = Y[lo:hi], g[lo:hi]
        sums = np.zeros((G, Y.shape[1]))
        np.add.at(sums, gl, Yl)
        counts = np.bincount(gl, minlength=G).astype(np.int64)
        meta = np.array([0, hi - lo], dtype=np.int64)
        allreduce(sums); allreduce(counts); allreduce(meta)
        ybar = sums.sum(axis=0) / meta[1]
        ms = min(Yl.shape[0], 64)                      # the library samples <= 4096 cells; 64 keeps the sample a strict subset here
        stride = Yl.shape[0] // ms if ms else 1
        psum = Yl[: ms * stride : stride].sum(axis=0) if ms else np.zeros(Y.shape[1])
        pcount = np.array([ms], dtype=np.int64)
        allreduce(psum); allreduce(pcount)
        a = psum / pcount[0]
        assert np.abs(a - ybar).max() > 1e-6            # really a different vector than the global mean
        gram = (Yl - a).T @ (Yl - a)
        allreduce(gram)
        mu = sums / counts[:, None]
        C = np.sqrt(counts)[:, None] * (mu - a)
        G0 = gram - C.T @ C
        Gc = gram - meta[1] * np.outer(ybar - a, ybar - a)
        # single-process truth
        full_sums = np.zeros((G, Y.shape[1]))
        np.add.at(full_sums, g, Y)
        Y0 = Y - (full_sums / np.bincount(g, minlength=G)[:, None])[g]
        ref = Y0.T @ Y0
        refc = (Y - Y.mean(axis=0)).T @ (Y - Y.mean(axis=0))
        q.put((rank, int(meta[1]), float(np.linalg.norm(G0 - ref) / np.linalg.norm(ref)), float(np.abs(G0).sum()),
               float(np.linalg.norm(Gc - refc) / np.linalg.norm(refc)), float(np.abs(a).sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_fit_decomposition_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [1500, 1500]           # the row counts were summed across ranks
    assert all(r[2] < 1e-12 for r in res)                  # sharded Gram of the residual == single process
    assert res[0][3] == res[1][3]                          # bit-identical on both ranks (no broadcast needed)
    assert all(r[4] < 1e-12 for r in res)                  # centred Gram of all cells recovered from the shifted one
    assert res[0][5] == res[1][5]                          # every rank agreed on the same shift vector


# ---- label tables across ranks (shards ordered by batch do not see every level) -------------------------------
class _FakeDev:
    def __init__(self, arr):
        self.arr = np.array(arr, copy=True)

    def to_host(self):
        return self.arr

    def free(self):
        pass


class _GlooCtx:
    """Stand-in for scmerge_b200.Context on CPU: same host-side protocol (to_device / allreduce / allgather_bytes), the sum
    all-reduce runs over gloo on host memory instead of NCCL on a device buffer."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def to_device(self, arr):
        return _FakeDev(arr)

    def allreduce(self, d):
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(d.arr)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def allgather_bytes(self, payload):
        from scmerge_b200.device import Context
        return Context.allgather_bytes(self, payload)


def _label_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from scmerge_b200.dist import shard_rows
    from scmerge_b200.ruv import _total_cells, replicate_keys
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = _GlooCtx(rank, world)
        rng = np.random.default_rng(4)
        m = 1001
        batch = np.sort(rng.choice(["b1", "b2", "b3"], size=m))                  # ordered by batch: rank 0 never sees "b3"
        types = rng.integers(0, 7, size=m) * 3                                     # integer labels with holes (0, 3, .., 18)
        types[:500][types[:500] == 18] = 0                                        # level 18 only occurs on the last rank
        pairs = np.stack([batch, types.astype(str)], axis=1)                      # data.frame-like labels
        lo, hi = shard_rows(m, world, rank)
        out = {}
        for name, lab, spare in (("batch", batch, None), ("types", types, None), ("types_spare", types, 10), ("pairs", pairs, None)):
            keys, G = replicate_keys(lab[lo:hi], spare=spare, ctx=ctx)
            ref_keys, ref_G = replicate_keys(lab, spare=spare)                     # single-process truth on ALL cells
            out[name] = (bool(np.array_equal(keys, ref_keys[lo:hi])), int(G), int(ref_G), int(len(set(lab[lo:hi].ravel().tolist()))))
        out["m_total"] = _total_cells(ctx, hi - lo)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_label_tables_are_global_across_ranks_gloo():
    import torch.multiprocessing as mp
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = _free_port()
    procs = [mpctx.Process(target=_label_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank in (0, 1):
        out = res[rank]
        assert out["m_total"] == 1001
        for name in ("batch", "types", "types_spare", "pairs"):
            same, G, ref_G, _seen = out[name]
            assert same and G == ref_G, (rank, name, out[name])                   # global ids and global G on every rank
    assert res[0]["batch"][1] == 3 and res[0]["batch"][3] < 3                      # rank 0 saw fewer batches than exist


def _pb_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from scmerge_b200.dist import shard_rows
    from scmerge_b200.scmerge2 import pseudobulk_keys
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = _GlooCtx(rank, world)
        rng = np.random.default_rng(8)
        m = 2003
        batch = np.sort(rng.choice(["b1", "b2", "b3"], size=m))                   # ordered by batch: rank 0 never sees "b3"
        clust = rng.choice(["ct_a", "ct_b", "ct_c", "ct_d"], size=m)
        clust[:1200][clust[:1200] == "ct_d"] = "ct_a"                             # "ct_d" only occurs on the last rank
        which = rng.integers(1, 8, size=m)                                        # a fixed fold draw (1..7) for all cells
        lo, hi = shard_rows(m, world, rank)
        keys, P, names, pb_t, pb_b = pseudobulk_keys(batch[lo:hi], clust[lo:hi], 7, which=which[lo:hi], ctx=ctx)
        rk, rP, rnames, rt, rb = pseudobulk_keys(batch, clust, 7, which=which)   # single-process truth on ALL cells
        drawn = pseudobulk_keys(batch[lo:hi], clust[lo:hi], 7, seed=3, ctx=ctx)   # folds drawn per rank: same table everywhere
        q.put((rank, dict(same=bool(np.array_equal(keys, rk[lo:hi])), P=P, rP=rP, names=names == rnames,
                          types=bool(np.array_equal(pb_t, rt)), batches=bool(np.array_equal(pb_b, rb)),
                          drawn_P=drawn[1], drawn_names=drawn[2], drawn_max=int(drawn[0].max()))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_pseudobulk_keys_are_global_across_ranks_gloo():
    """scMerge2 on several ranks: the (batch, fold, cluster) table, the keys and the row names of the pseudobulk matrix must be
    those of the single-process construction even when a shard lacks a batch or a cluster."""
    import torch.multiprocessing as mp
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = _free_port()
    procs = [mpctx.Process(target=_pb_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank in (0, 1):
        out = res[rank]
        assert out["same"] and out["P"] == out["rP"] and out["names"] and out["types"] and out["batches"], (rank, out)
    assert res[0]["drawn_P"] == res[1]["drawn_P"] and res[0]["drawn_names"] == res[1]["drawn_names"]
    assert max(res[0]["drawn_max"], res[1]["drawn_max"]) == res[0]["drawn_P"] - 1

"""torchrun worker for the multi-GPU parity test: every rank holds a contiguous block of cells, the fit runs with the library's
own NCCL communicator (SCM_TEST_TRANSPORT=hook: the torch.distributed hook instead), rank-local adjust, and each rank checks its
rows against the single-process oracle.  The cells are ORDERED BY BATCH and one cell type only occurs in the last rows, so the
shards do not see every label: the level tables must be built over all ranks (scmerge_b200.ruv._global_levels)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)  # a rank stuck in a collective prints where, instead of hanging the test
    import torch
    import torch.distributed as dist

    import scmerge_b200 as sm
    from oracle import ruv_oracle as orc
    from scmerge_b200.dist import init_builtin_nccl, shard_rows, torch_allreduce_hook
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    transport = os.environ.get("SCM_TEST_TRANSPORT", "nccl")
    if transport == "hook":
        ctx = sm.Context(local, stream=torch.cuda.current_stream().cuda_stream)
        ctx.set_allreduce(torch_allreduce_hook(local), world, rank)
    else:
        ctx = init_builtin_nccl(sm.Context(local), rank, world)
    assert ctx.transport() == transport
    if os.environ.get("SCM_TEST_DUMP"):
        # bit-for-bit twin of tests/test_gpu_multi.py::test_one_process_two_gpus_matches_the_torchrun_path: the same host matrix,
        # the same balanced row blocks, the per-rank host driver under torchrun; every rank dumps its rows
        Pd = orc.planted_factors(9001, 384, 128, 10, n_types=6, n_batch=3, seed=31)
        base, rem = 9001 // world, 9001 % world
        lo = rank * base + min(rank, rem)
        hi = lo + base + (1 if rank < rem else 0)
        out = sm.fastRUVIII(np.ascontiguousarray(Pd["Y"][lo:hi]), Pd["types"][lo:hi], Pd["ctl"], k=10, return_info=True, ctx=ctx)
        np.savez(os.environ["SCM_TEST_DUMP"] + f".rank{rank}.npz", newY=out["newY"], fullalpha=np.asarray(out["fullalpha"]), lo=lo, hi=hi)
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0)
    P = orc.planted_factors(9000, 384, 128, 10, n_types=6, n_batch=3, seed=31)
    # batch-ordered cells (stable), labels as strings; cell type "t5" is moved to the very end: rank 0 never sees it
    order = np.argsort(P["batch"], kind="stable")
    order = np.concatenate([order[P["types"][order] != 5], order[P["types"][order] == 5]])
    Yall, types_all = P["Y"][order], np.array([f"t{t}" for t in P["types"][order]])
    batch_all = np.array([f"b{b}" for b in P["batch"][order]])
    lo, hi = shard_rows(9000, world, rank)
    if rank == 0 and world > 1:  # the first shard must really lack a batch and a cell type
        assert "b2" not in set(batch_all[lo:hi]) and "t5" not in set(types_all[lo:hi]), "shard 0 is supposed to miss levels"
    # fastRUVIII path
    out = sm.fastRUVIII(Yall[lo:hi], types_all[lo:hi], P["ctl"], k=10, return_info=True, ctx=ctx)
    ref = orc.fastRUVIII(Yall, types_all, P["ctl"], 10, return_info=True)
    e1 = orc.rel_frobenius(out["newY"], ref["newY"][lo:hi])
    a1 = float(orc.principal_angles(out["fullalpha"][:10], ref["fullalpha"][:10]).max())
    # device-resident shard (scm_ruv3_fit + scm_adjust instead of the host pipeline)
    dY = sm.DeviceMatrix.from_host(ctx, Yall[lo:hi])
    outd = sm.fastRUVIII(dY, types_all[lo:hi], P["ctl"], k=10, return_info=True, ctx=ctx)
    e1d = orc.rel_frobenius(outd["newY"].to_host(), ref["newY"][lo:hi])
    dY.free()
    # scRUVIII path (standardisation statistics also cross ranks)
    res = sm.scRUVIII(Yall[lo:hi].T, types_all[lo:hi], P["ctl"], k=10, batch=batch_all[lo:hi], ctx=ctx)
    ref2 = orc.scRUVIII(Yall.T, types_all, P["ctl"], 10, batch_all)
    e2 = orc.rel_frobenius(res[10]["newY"], ref2[10]["newY"][lo:hi])
    # ruvK selection across ranks (scores and labels gathered, distance pass split over the ranks) and scRUVg
    sel = sm.scRUVIII(Yall[lo:hi].T, types_all[lo:hi], P["ctl"], k=[4, 10], batch=batch_all[lo:hi], cell_type=types_all[lo:hi], ctx=ctx)
    rsel = orc.scRUVIII(Yall.T, types_all, P["ctl"], [4, 10], batch_all, cell_type=types_all)
    e3 = float(np.abs(sel["sil"] - rsel["sil"]).max())
    same_k = sel["optimal_ruvK"] == rsel["optimal_ruvK"]
    e4 = orc.rel_frobenius(sel[4]["newY"], rsel[4]["newY"][lo:hi])
    rg = sm.scRUVg(Yall[lo:hi], P["ctl"], 8, ctx=ctx)
    # reference: the definition (first k left singular vectors of the centred control block, alpha = W'Y0) without the
    # oracle's literal m x m cross product of the >= 5:1 branch (9000 x 9000 here; equality of the two is a CPU test)
    Y0 = Yall - Yall.mean(axis=0)
    U = np.linalg.svd(Y0[:, orc.tological(P["ctl"], 384)], full_matrices=False)[0][:, :8]
    e5 = orc.rel_frobenius(rg["newY"], (Yall - U @ (U.T @ Y0))[lo:hi])
    # scRUVg with a GIVEN W sharded by rows (W'W, W'Y0 and W'1 are all-reduced in scm_ls_coef), and eta (ruv::RUV1: rank-local given
    # the covariates) in front of the cross-rank fit
    rgw = sm.scRUVg(Yall[lo:hi], P["ctl"], 5, fullW=U[lo:hi], ctx=ctx)
    W5 = U[:, :5]
    e5w = orc.rel_frobenius(rgw["newY"], (Yall - W5 @ np.linalg.solve(W5.T @ W5, W5.T @ Y0))[lo:hi])
    eta = np.random.default_rng(8).normal(size=(384, 2))
    oeta = sm.fastRUVIII(Yall[lo:hi], types_all[lo:hi], P["ctl"], k=6, eta=eta, ctx=ctx)
    e5e = orc.rel_frobenius(oeta, orc.fastRUVIII(Yall, types_all, P["ctl"], 6, eta=eta)[lo:hi])
    # scMerge2 core across ranks (dense and sparse input): pseudobulks pool cells of ALL ranks (global level tables, sums and
    # counts all-reduced before the division), the pseudobulk fit is shared row-wise, the adjust is rank-local.  The fold of
    # every cell is drawn once for all cells and handed out with the shards, so the pseudobulks are those of the oracle.
    from scmerge_b200.scmerge2 import scMerge2_core
    import scipy.sparse as sp
    Ypos = np.abs(Yall) + 0.1
    rng = np.random.default_rng(5)
    which_all = np.empty(9000, dtype=np.int64)
    for bname in np.unique(batch_all):
        idx = np.nonzero(batch_all == bname)[0]
        w = np.empty(idx.shape[0], dtype=np.int64)
        w[rng.permutation(idx.shape[0])] = (np.arange(idx.shape[0]) % 10) + 1
        which_all[idx] = w
    ref6 = orc.scmerge2_core(Ypos.T, batch_all, types_all, which_all, ruvK=6, k_fold=10, exact=True)
    o6 = scMerge2_core(Ypos[lo:hi], batch_all[lo:hi], types_all[lo:hi], ruvK=6, k_pseudoBulk=10, which=which_all[lo:hi],
                       BSPARAM=sm.ExactParam(), ctx=ctx)
    e6 = orc.rel_frobenius(o6["newY"], ref6["newY"][lo:hi])
    names_ok = o6["pseudobulk_names"] == ref6["names"]
    e6b = float(np.abs(o6["bulkExprs"] - ref6["bulkExprs"]).max())
    mask = np.random.default_rng(6).random(Ypos.shape) < 0.3
    Ysp = np.where(mask, Ypos, 0.0)
    ref7 = orc.scmerge2_core(Ysp.T, batch_all, types_all, which_all, ruvK=6, k_fold=10, exact=True)
    o7 = scMerge2_core(sp.csr_matrix(Ysp[lo:hi]), batch_all[lo:hi], types_all[lo:hi], ruvK=6, k_pseudoBulk=10, which=which_all[lo:hi],
                       BSPARAM=sm.ExactParam(), ctx=ctx)
    e7 = orc.rel_frobenius(o7["newY"], ref7["newY"][lo:hi])
    # every rank must hold bit-identical fullalpha (no broadcast in the design)
    fa = torch.from_numpy(np.array(out["fullalpha"])).cuda()
    lst = [torch.empty_like(fa) for _ in range(world)]
    dist.all_gather(lst, fa)
    same = all(torch.equal(lst[0], t) for t in lst)
    ok = (e1 < 1e-8 and e1d < 1e-8 and a1 < 1e-6 and e2 < 1e-8 and same and e3 < 1e-7 and same_k and e4 < 1e-8 and e5 < 1e-8 and e5w < 1e-8 and e5e < 1e-8
          and e6 < 1e-8 and names_ok and e6b < 1e-12 and e7 < 1e-8)
    print(f"rank {rank}/{world} [{transport}]: rows [{lo},{hi}) batches here {sorted(set(batch_all[lo:hi]))} fast err {e1:.2e} (device {e1d:.2e}) "
          f"angle {a1:.2e} scRUVIII err {e2:.2e} identical_alpha {same} selection sil err {e3:.2e} same_k {same_k} newY err {e4:.2e} "
          f"scRUVg err {e5:.2e} (fullW {e5w:.2e}) eta err {e5e:.2e} scMerge2 err {e6:.2e} (names {names_ok}, bulk {e6b:.1e}) sparse {e7:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

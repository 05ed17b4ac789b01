"""torchrun worker for the multi-GPU parity test: every rank holds a contiguous block of cells, the fit runs with
the NCCL all-reduce hook, rank-local adjust, and each rank checks its rows against the single-process oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import scmerge_b200 as sm
    from oracle import ruv_oracle as orc
    from scmerge_b200.dist import shard_rows, torch_allreduce_hook
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sm.Context(local, stream=torch.cuda.current_stream().cuda_stream)
    ctx.set_allreduce(torch_allreduce_hook(local), world)
    P = orc.planted_factors(9000, 384, 128, 10, n_types=6, n_batch=3, seed=31)
    lo, hi = shard_rows(9000, world, rank)
    ok = True
    # fastRUVIII path
    out = sm.fastRUVIII(P["Y"][lo:hi], P["types"][lo:hi], P["ctl"], k=10, return_info=True, ctx=ctx)
    ref = orc.fastRUVIII(P["Y"], P["types"], P["ctl"], 10, return_info=True)
    e1 = orc.rel_frobenius(out["newY"], ref["newY"][lo:hi])
    a1 = float(orc.principal_angles(out["fullalpha"][:10], ref["fullalpha"][:10]).max())
    # scRUVIII path (standardisation statistics also cross ranks)
    batch = np.array([f"b{b}" for b in P["batch"]])
    res = sm.scRUVIII(P["Y"][lo:hi].T, P["types"][lo:hi], P["ctl"], k=10, batch=batch[lo:hi], ctx=ctx)
    ref2 = orc.scRUVIII(P["Y"].T, P["types"], P["ctl"], 10, batch)
    e2 = orc.rel_frobenius(res[10]["newY"], ref2[10]["newY"][lo:hi])
    # every rank must hold bit-identical fullalpha (no broadcast in the design)
    fa = torch.from_numpy(out["fullalpha"].copy()).cuda()
    lst = [torch.empty_like(fa) for _ in range(world)]
    dist.all_gather(lst, fa)
    same = all(torch.equal(lst[0], t) for t in lst)
    ok = e1 < 1e-8 and a1 < 1e-6 and e2 < 1e-8 and same
    print(f"rank {rank}/{world}: rows [{lo},{hi}) fast err {e1:.2e} angle {a1:.2e} scRUVIII err {e2:.2e} identical_alpha {same}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

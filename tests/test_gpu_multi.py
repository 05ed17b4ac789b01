"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise).
  * one process per GPU (torchrun): cells sharded over 2 ranks, the library's own NCCL communicator (and, second run, the
    torch.distributed hook), compared with the single-process oracle; shards that do not see every batch / cell type,
  * ONE process driving 2 GPUs (scm_multi_create: what a single R session uses), bit-for-bit against the torchrun path's
    arithmetic (same shards, same collectives) and against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("transport", ["nccl", "hook"])
def test_two_gpu_sharded_fit_and_adjust(transport):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531" if transport == "nccl" else "29532", os.path.join(ROOT, "tests", "multi_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, SCM_TEST_TRANSPORT=transport))
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0


def test_one_process_two_gpus_matches_oracle_and_is_deterministic():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import scmerge_b200 as sm
    from oracle import ruv_oracle as orc
    P = orc.planted_factors(9001, 384, 128, 10, n_types=6, n_batch=3, seed=31)  # odd cell count: unequal shards
    mg = sm.MultiContext(devices=[0, 1])
    assert mg.size == 2 and mg.ctx(1).transport() == "nccl"
    out = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=10, return_info=True, ctx=mg)
    ref = orc.fastRUVIII(P["Y"], P["types"], P["ctl"], 10, return_info=True)
    assert orc.rel_frobenius(out["newY"], ref["newY"]) < 1e-8
    assert orc.principal_angles(out["fullalpha"][:10], ref["fullalpha"][:10]).max() < 1e-6
    out2 = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=10, return_info=True, ctx=mg)
    np.testing.assert_array_equal(out["newY"], out2["newY"])            # run-to-run bit identical
    batch = np.array([f"b{b}" for b in P["batch"]])
    res = sm.scRUVIII(np.asfortranarray(P["Y"].T), P["types"], P["ctl"], k=10, batch=batch, ctx=mg)
    ref2 = orc.scRUVIII(P["Y"].T, P["types"], P["ctl"], 10, batch)
    assert orc.rel_frobenius(res[10]["newY"], ref2[10]["newY"]) < 1e-8


def test_one_process_two_gpus_matches_the_torchrun_path(tmp_path):
    """ONE process driving 2 GPUs (scm_multi_ruv3_host: ncclCommInitAll, a host thread per device) against one process per GPU
    (torchrun, scm_ctx_comm_init_nccl) on the same host matrix and the same row blocks: same kernels, same collectives, so the
    adjusted matrix and fullalpha must agree BIT FOR BIT."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import scmerge_b200 as sm
    from oracle import ruv_oracle as orc
    dump = str(tmp_path / "torchrun")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, SCM_TEST_TRANSPORT="nccl", SCM_TEST_DUMP=dump))
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
    parts = [np.load(dump + f".rank{i}.npz") for i in range(2)]
    assert int(parts[0]["lo"]) == 0 and int(parts[0]["hi"]) == int(parts[1]["lo"]) and int(parts[1]["hi"]) == 9001
    P = orc.planted_factors(9001, 384, 128, 10, n_types=6, n_batch=3, seed=31)
    mg = sm.MultiContext(devices=[0, 1])
    out = sm.fastRUVIII(P["Y"], P["types"], P["ctl"], k=10, return_info=True, ctx=mg)
    np.testing.assert_array_equal(np.concatenate([parts[0]["newY"], parts[1]["newY"]]), out["newY"])
    np.testing.assert_array_equal(parts[0]["fullalpha"], np.asarray(out["fullalpha"]))
    np.testing.assert_array_equal(parts[1]["fullalpha"], np.asarray(out["fullalpha"]))

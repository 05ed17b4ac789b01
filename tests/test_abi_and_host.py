"""CPU tests (no GPU needed): the C-ABI library loads and exports every symbol declared in include/scmerge_b200.h,
the product path fails loudly without a device, and the host-side argument logic mirrors the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "scmerge_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scm_[a-z0-9_]+)\s*\(", text)) - {"scm_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    from scmerge_b200 import _lib
    handle = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in the header but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert handle.scm_version() >= 100
    assert sorted(_lib.SIGNATURES) == syms


def test_no_cpu_fallback_without_device():
    import scmerge_b200 as sm
    h = C.c_void_p()
    rc = sm.lib().scm_ctx_create(0, None, C.byref(h))
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        assert rc == 0
        sm.lib().scm_ctx_destroy(h)
    else:
        assert rc == sm._lib.ERR_CUDA
        assert b"no CPU fallback" in sm.lib().scm_last_error()
        with pytest.raises(sm.ScmError):
            sm.fastRUVIII(np.zeros((8, 4)), np.arange(8) % 2, np.arange(4), k=1)


def test_product_path_never_imports_the_oracle():
    for dirpath, _d, files in os.walk(os.path.join(ROOT, "scmerge_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"
                assert "/root/reference" not in src


def test_replicate_keys_and_matrix():
    from scmerge_b200 import replicate_keys, replicate_matrix
    keys, G = replicate_keys(np.array(["b", "a", "b", "c"]))
    assert G == 3 and keys.tolist() == [1, 0, 1, 2] and keys.dtype == np.int32  # sorted level order
    M = replicate_matrix(np.array(["b", "a", "b", "c"]))
    assert M.tolist() == [[0, 1, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]
    keys2, G2 = replicate_keys(M)
    assert G2 == 3 and np.array_equal(keys2, keys)
    keys3, G3 = replicate_keys(np.array([["t1", "b1"], ["t1", "b2"], ["t1", "b1"]]))  # data.frame -> pasted
    assert G3 == 2 and keys3.tolist() == [0, 1, 0]
    keys4, G4 = replicate_keys(np.array([3, 3, 7, 5], dtype=np.int64))  # integer labels
    assert G4 == 3 and keys4.tolist() == [0, 0, 2, 1]
    # `spare`: with enough slack the histogram that drops unused label values is skipped (empty groups are harmless) ...
    lab = np.array([0, 2, 2, 5] * 50, dtype=np.int32)
    k5, G5 = replicate_keys(lab, spare=100)
    assert G5 == 6 and k5 is lab
    # ... but not when m - G could enter the svd_k rule: then the levels are compacted like R's factor()
    k6, G6 = replicate_keys(lab, spare=198)
    assert G6 == 3 and k6.tolist()[:4] == [0, 1, 1, 2]
    with pytest.raises(ValueError, match="one-hot"):
        replicate_keys(np.array([[1.0, 1.0], [0.0, 1.0]]))
    with pytest.raises(ValueError, match="one-hot"):
        replicate_keys(np.array([[0.5, 0.5], [0.0, 1.0]]))


def test_tological_chunks_and_svd_k_rule():
    from scmerge_b200 import createChunk, tological
    from scmerge_b200.ruv import _device_rows_cap, _svd_k
    assert tological([0, 2], 4).tolist() == [True, False, True, False]
    assert tological(np.array([True, False]), 3).tolist() == [True, False, False]
    assert createChunk(9, 4) == [(1, 4), (5, 9)] and createChunk(10, 4) == [(1, 4), (5, 8), (9, 10)]
    assert createChunk(120000, 50000) == [(1, 50000), (50001, 120000)]
    # R/fastRUVIII.R:66-70: exact caps at svd_k, non-exact ignores it (quirk)
    assert _svd_k(400, 6, 400, True, 50) == 50 and _svd_k(400, 6, 400, False, 50) == 394
    assert _svd_k(40, 6, 400, True, 50) == 34
    assert _device_rows_cap(50, 20, True) == (50, True)
    # more rows than a subspace block carries: the whole spectrum when the Gram is small enough, else truncation with a warning
    assert _device_rows_cap(394, 20, False, 2000) == (394, True)
    with pytest.warns(UserWarning, match="truncated"):
        assert _device_rows_cap(394, 20, False, 20000) == (50, True)
    assert _device_rows_cap(40, 20, False) == (40, False)


def test_early_out_needs_no_device():
    import scmerge_b200 as sm
    Y = np.arange(12.0).reshape(6, 2)
    out = sm.fastRUVIII(Y, np.arange(6), [0], k=3, return_info=True)  # ncol(M) >= m (R/fastRUVIII.R:78-80)
    assert out["newY"] is Y and out["fullalpha"] is None
    assert sm.fastRUVIII(Y, np.arange(6) % 2, [0], k=0) is Y
    with pytest.raises(TypeError):
        sm.fastRUVIII(Y, np.arange(6) % 2, [0])
    with pytest.raises(ValueError):
        sm.fastRUVIII(Y, np.arange(6) % 2, [0], k=1, eta=np.ones((3, 5)))  # eta must have one row / column per gene: device-free error


def test_scruviii_early_out_needs_no_device():
    """scRUVIII on a degenerate replicate structure: fastRUVIII returns its input (R/fastRUVIII.R:78-80) and the
    de-standardisation (R/scRUVIII.R:117-119) gives the original matrix back, cells x genes, fullalpha NULL."""
    import scmerge_b200 as sm
    Y = np.arange(12.0).reshape(2, 6)  # genes x cells
    res = sm.scRUVIII(Y, np.arange(6), [0], k=3, batch=np.arange(6) % 2)
    assert res["optimal_ruvK"] == 3 and res[3]["fullalpha"] is None
    np.testing.assert_array_equal(res[3]["newY"], Y.T)
    one = sm.scRUVIII(Y, np.arange(6) % 2, [0], k=[0, 2], batch=np.arange(6) % 2, return_all_RUV=False)
    assert one["optimal_ruvK"] == 0 and one["fullalpha"] is None and one["newY"].shape == (6, 2)
    with pytest.raises(ValueError, match="one row per cell"):
        sm.scRUVIII(Y, np.arange(5), [0], k=1, batch=np.arange(6) % 2)


def test_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-cells", "1500", "--genes", "300", "--ctl", "100"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "RUV-III cells/sec" and d["unit"] == "cells/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def _call_sites(text):
    """(.Call name, number of arguments) for every `.Call("name", ...)` in R source (comments stripped)."""
    text = "\n".join(ln.split("#")[0] for ln in text.splitlines())
    out = []
    for mt in re.finditer(r'\.Call\(\s*"([A-Za-z0-9_]+)"', text):
        i, depth, nargs = mt.end(), 1, 0
        while depth:
            ch = text[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                nargs += 1
            i += 1
        out.append((mt.group(1), nargs))
    return out


def test_r_shim_typechecks_and_matches_the_r_closures():
    """The R side cannot be built or run here (no R).  What CAN be checked on CPU: (1) r-package/src/rshim.c compiles
    (syntax + types) against include/scmerge_b200.h and a stand-in for R's C API (tests/r_stub), so every C-ABI call in it has
    the declared arity and pointer types; (2) every `.Call("name", ...)` in r-package/R/*.R is registered in the shim's
    R_CallMethodDef table with that number of arguments."""
    import glob
    import shutil
    import subprocess
    shim = os.path.join(ROOT, "r-package", "src", "rshim.c")
    if shutil.which("gcc"):
        r = subprocess.run(["gcc", "-fsyntax-only", "-std=c11", "-Wall", "-Werror=implicit-function-declaration",
                            "-Werror=incompatible-pointer-types", "-Werror=int-conversion", "-Wno-cast-function-type",
                            "-I" + os.path.join(ROOT, "tests", "r_stub"), "-I" + os.path.join(ROOT, "include"), shim],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    table = dict((nm, int(na)) for nm, na in re.findall(r'\{"([a-z0-9_]+)",\s*\(DL_FUNC\)&[a-z0-9_]+,\s*(\d+)\}', open(shim).read()))
    assert table, "no R_CallMethodDef entries found"
    sites = []
    for path in glob.glob(os.path.join(ROOT, "r-package", "R", "*.R")):
        sites += _call_sites(open(path).read())
    assert len(sites) >= 10
    for name, nargs in sites:
        assert name in table, f'.Call("{name}") is not registered in rshim.c'
        assert table[name] == nargs, f'.Call("{name}") passes {nargs} arguments, the shim registers {table[name]}'


def test_pseudobulk_keys_match_the_oracle_bit_exactly():
    """Host logic only (no device): the composite (batch, fold, cluster) key of every cell reproduces the row order, the
    names and the group membership of the reference's `bulkExprs` (create_pseudoBulk batch by batch,
    R/scMerge2_utilis.R:259-263,461-486), including droplevels() of empty combinations and ragged batches (a batch with
    fewer cells than k_fold, a cluster missing from a batch)."""
    import sys
    sys.path.insert(0, ROOT)
    from oracle import ruv_oracle as orc
    from scmerge_b200.scmerge2 import pseudobulk_keys
    rng = np.random.default_rng(11)
    m, n, K = 137, 9, 6
    batch = np.array(["b2"] * 70 + ["b10"] * 63 + ["b3"] * 4)            # "b10" < "b2" < "b3": sorted LEVEL order, 4 cells < K
    clust = rng.choice(["T", "B", "NK"], size=m)
    clust[batch == "b3"] = "T"                                          # clusters B / NK are missing from batch b3
    X = rng.normal(size=(n, m))                                         # genes x cells
    which = np.empty(m, dtype=np.int64)
    for b in np.unique(batch):
        idx = np.nonzero(batch == b)[0]
        Kb = min(idx.shape[0], K)
        which[idx[rng.permutation(idx.shape[0])]] = (np.arange(idx.shape[0]) % Kb) + 1   # cvFolds: rep(1:K, length.out) over a random order
    keys, P, names, pb_t, pb_b = pseudobulk_keys(batch, clust, K, which=which)
    bulks, ref_names = [], []
    for i, b in enumerate(np.unique(batch), start=1):
        sel = batch == b
        bulk, nm = orc.create_pseudobulk(X[:, sel], clust[sel], which[sel], k_fold=min(int(sel.sum()), K))
        bulks.append(bulk)
        ref_names += [f"{i}_{x}" for x in nm]                         # paste(i, rownames(res)): the batch INDEX (R/scMerge2_utilis.R:266)
    ref = np.concatenate(bulks)
    assert names == ref_names and P == ref.shape[0] == len(set(names))
    assert keys.dtype == np.int32 and keys.min() == 0 and keys.max() == P - 1
    for p in range(P):                                                  # same cells in every pseudobulk (bit-exact indices)
        cells = np.nonzero(keys == p)[0]
        np.testing.assert_array_equal(X[:, cells].sum(axis=1) / cells.shape[0], ref[p])
    tlev, blev = np.unique(clust), np.unique(batch)
    assert [f"{b + 1}_{tlev[t]}" for b, t in zip(pb_b, pb_t)] == [x.rsplit("_", 1)[0] for x in names]
    assert names[0].split("_")[0] == "1" and list(blev) == ["b10", "b2", "b3"]          # index 1 = first SORTED level ("b10")
    # a seeded draw has the cvFolds structure: fold sizes within a batch differ by at most one
    k2, P2, *_ = pseudobulk_keys(batch, clust, K, seed=1)
    assert k2.shape == (m,) and 0 < P2 <= 3 * 3 * K


def test_create_chunk_and_replicate_keys_against_independent_restatements():
    """Two independently written versions of the same reference logic must agree everywhere: createChunk
    (R/scMerge2_utilis.R:153-170) of the host mirror vs the oracle, and replicate_keys vs the oracle's one-hot
    ruv::replicate.matrix for label vectors, data-frame-like labels and one-hot matrices."""
    import sys
    sys.path.insert(0, ROOT)
    from oracle import ruv_oracle as orc
    from scmerge_b200.ruv import createChunk, replicate_keys, replicate_matrix
    for n in list(range(1, 60)) + [999, 1000, 1001, 1499, 1500, 1501, 50_000, 74_999, 75_000, 125_001]:
        for cs in (1, 2, 7, 10, 1000, 50_000):
            assert createChunk(n, cs) == orc.create_chunk(n, cs), (n, cs)
    rng = np.random.default_rng(2)
    for labels in (rng.integers(0, 5, 40), rng.integers(3, 90, 200) * 7, np.array(list("abcabcddz")), rng.normal(size=12).round(1),
                   np.stack([rng.integers(0, 3, 30), rng.integers(0, 2, 30)], axis=1).astype(str)):
        keys, G = replicate_keys(labels)
        M_ref = orc.replicate_matrix(labels)
        assert M_ref.shape == (len(labels), G)
        np.testing.assert_array_equal(np.argmax(M_ref, axis=1), keys)
        np.testing.assert_array_equal(replicate_matrix(labels), M_ref)
        k2, G2 = replicate_keys(M_ref)                       # a one-hot matrix passes through with its column order
        assert G2 == G
        np.testing.assert_array_equal(k2, keys)
    with pytest.raises(ValueError, match="one-hot"):
        replicate_keys(np.array([[1.0, 1.0], [0.0, 1.0]]))


def test_every_entry_point_rejects_null_arguments_without_crashing():
    """No exception or signal may cross the C ABI: every exported function called with a NULL context / NULL pointers / zero
    sizes returns a negative status (scm_ctx_destroy(NULL) is a no-op) and leaves a message in scm_last_error().  Run in a
    subprocess so that a segfault would be reported as a failure instead of killing the test session."""
    import subprocess
    import sys
    code = """
import ctypes as C, sys
sys.path.insert(0, %r)
from scmerge_b200 import _lib as L
lib = L.lib()
for name, args in L.SIGNATURES.items():
    vals = [L.ALLREDUCE_FN(0) if a is L.ALLREDUCE_FN else 0.0 if a is C.c_double else None if (a in (C.c_void_p, C.c_char_p) or hasattr(a, "contents")) else 0
            for a in args]
    rc = getattr(lib, name)(*vals)
    if name in ("scm_last_error", "scm_version"):
        continue
    if name == "scm_multi_ctx":  # returns a pointer: NULL for a NULL handle
        assert rc is None, rc
        continue
    ok = rc == 0 if name in ("scm_ctx_destroy", "scm_multi_destroy", "scm_multi_size") else rc <= 0 if name == "scm_host_free_pinned" else rc < 0  # cudaFreeHost(NULL) succeeds on a GPU box
    assert ok, (name, rc)
    if rc < 0:
        assert lib.scm_last_error(), name
print("OK")
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), (r.returncode, r.stdout[-500:], r.stderr[-2000:])


def test_shard_rows_partition():
    from scmerge_b200.dist import shard_rows
    for m, w in [(10, 3), (7, 8), (1_000_000, 8), (0, 2)]:
        parts = [shard_rows(m, w, r) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == m
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))


def test_gene_selection_by_name():
    """getAdjustedMat / scMerge2 select control and subset genes by NAME (R/scMerge2.R:304,382-390): unknown names are
    dropped, the matrix order is kept."""
    from scmerge_b200 import genes_by_name
    names = np.array(["g3", "g1", "g7", "g2"])
    assert genes_by_name(["g2", "g3", "nope"], 4, names).tolist() == [True, False, False, True]
    assert genes_by_name(np.array([1, 2]), 4).tolist() == [False, True, True, False]
    assert genes_by_name(np.array([True, False, True, False]), 4).tolist() == [True, False, True, False]
    with pytest.raises(ValueError, match="gene_names"):
        genes_by_name(["g1"], 4)


def test_eta_design_rows_match_the_oracle_design_matrix():
    """Host-side resolution of gene-wise covariates (ruv::RUV1's design.matrix) against the oracle's restatement."""
    from scmerge_b200.ruv import eta_design_rows
    from oracle import ruv_oracle as orc
    n = 40
    rng = np.random.default_rng(0)
    eta = rng.normal(size=(n, 2))
    for arg, inter in ((eta, True), (eta, False), (eta.T, True), (eta[:, 0], True), (1, True), (np.c_[np.ones(n), eta], True)):
        assert np.array_equal(eta_design_rows(arg, n, inter), orc.design_matrix(arg, n, inter).T)
    with pytest.raises(ValueError):
        eta_design_rows(np.c_[eta, eta[:, :1] * 2], n, True)                # collinear
    with pytest.raises(ValueError):
        eta_design_rows(np.zeros((n + 1, 2)), n, True)                      # wrong shape
    with pytest.raises(ValueError):
        eta_design_rows(2, n, True)


def test_committed_ncu_traffic_stamp_matches_the_kernel_sources():
    """bench.py attaches `roofline.traffic` (DRAM bytes per launch from an ncu capture) only when profiles/ncu_traffic.json was
    captured for the kernel sources in the tree.  A kernel edit without a new capture (tools/probe/ncu_traffic.sh on a B200)
    would silently turn the figure into null: fail here instead."""
    import json
    import bench
    stamp = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert stamp["kernel_source_hash"] == bench.kernel_source_hash(), \
        "kernel sources changed since profiles/ncu_traffic.json was captured: re-run tools/probe/ncu_traffic.sh under gpurun"
    algo = stamp["algorithmic_bytes"]
    for kern, got in stamp["dram_bytes_per_launch"].items():
        assert got >= 0.99 * algo[kern], kern                       # never below the algorithmic bytes
    assert stamp["dram_bytes_per_launch"]["gram_tma_kernel"] <= 4.0 * algo["gram_tma_kernel"]   # <= 4x


def test_host_argument_resolution_properties():
    """Property checks (hypothesis) of the host-side index plumbing against the oracle's restatements: replicate keys are a
    relabelling of the input partition with levels in sorted order, createChunk covers 1..n without gaps like the reference's,
    the svd_k rule, and `subset` resolution keeps the given order."""
    from hypothesis import given, settings, strategies as st
    from scmerge_b200.ruv import _resolve_subset, _svd_k, createChunk, replicate_keys, replicate_matrix
    from oracle import ruv_oracle as orc

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(min_value=-5, max_value=40), min_size=1, max_size=200))
    def keys_ok(labels):
        a = np.array(labels)
        k, G = replicate_keys(a)
        lv = np.unique(a)
        assert G == lv.shape[0] and np.array_equal(lv[k], a)                    # sorted level order, a pure relabelling
        assert np.array_equal(replicate_matrix(a), orc.replicate_matrix(a))
        ks, Gs = replicate_keys(np.array([f"t{v}" for v in labels]))            # strings sort like R's factor levels (C locale)
        assert Gs == G and len(set(zip(ks.tolist(), k.tolist()))) == G

    @settings(max_examples=60, deadline=None)
    @given(st.integers(min_value=1, max_value=5000), st.integers(min_value=1, max_value=700))
    def chunks_ok(n, size):
        ch = createChunk(n, size)
        assert ch == orc.create_chunk(n, size)
        assert ch[0][0] == 1 and ch[-1][1] == n and all(b[0] == a[1] + 1 for a, b in zip(ch, ch[1:]))

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 10**6), st.integers(1, 500), st.integers(1, 5000), st.booleans(), st.integers(1, 200))
    def svdk_ok(m, G, nctl, exact, cap):
        assert _svd_k(m, G, nctl, exact, cap) == orc.svd_k_bound(m, G, nctl, exact, cap)

    @settings(max_examples=40, deadline=None)
    @given(st.lists(st.integers(0, 29), min_size=1, max_size=40))
    def subset_ok(idx):
        names = np.array([f"g{i}" for i in range(30)])
        assert np.array_equal(_resolve_subset(np.array(idx), 30), idx)                              # order and repeats kept
        assert np.array_equal(_resolve_subset(names[idx], 30, gene_names=names), idx)               # match(subset, colnames(Y))
        mask = np.zeros(30, bool)
        mask[idx] = True
        assert np.array_equal(_resolve_subset(mask, 30), np.nonzero(mask)[0])                       # which()

    keys_ok(); chunks_ok(); svdk_ok(); subset_ok()
    with pytest.raises(IndexError):
        _resolve_subset(np.array(["nope"]), 30, gene_names=np.array([f"g{i}" for i in range(30)]))

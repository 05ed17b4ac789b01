/* .Call shim between R and libscmerge_b200 (include/scmerge_b200.h).  Logic-free by design: it unpacks
 * SEXPs, forwards to the C ABI and turns non-zero status codes into R errors.  NOT compiled in this
 * repository's environment (no R headers here; type-checked against tests/r_stub); see INTEGRATION.md for the build line.
 *
 * Replaces, inside scMerge's own closures, the arithmetic of
 *   fastRUVIII  (R/fastRUVIII.R:41-111)      -> scm_fit(), scm_adjust_()
 *   scRUVIII    (R/scRUVIII.R:41-137)        -> scm_fit() with batch keys, scm_adjust_() with center/pre/post
 *   pseudoRUVIII (R/scMerge2_utilis.R:17-150), getAdjustedMat (R/scMerge2.R:368-410) -> same two calls
 *   create_pseudoBulk / aggregate.Matrix (R/scMerge2_utilis.R:325-346,461-486)       -> scm_segmean()
 *   as(., "CsparseMatrix") + batchelor::cosineNorm (R/scMerge2.R:164-166,193-198)     -> scm_upload_sparse()
 *   scRUVg (R/scRUVg.R:30-85)                                                         -> scm_ruvg()
 *   calculateSil / kBET_batch_sil (R/scRUVIII.R:182-199)                              -> scm_fit_gram() + scm_sil()
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include "scmerge_b200.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != 0 && rc_ != SCM_ERR_NOCONV) Rf_error("%s", scm_last_error()); \
                         if (rc_ == SCM_ERR_NOCONV) Rf_warning("%s", scm_last_error()); } while (0)

static scm_ctx* g_ctx = NULL;
static scm_multi* g_multi = NULL;   /* set by scm_use_gpus(): ONE R process driving several GPUs (scm_multi_create) */
static scm_ctx* ctx(void) {
    if (g_multi) return scm_multi_ctx(g_multi, 0);
    if (!g_ctx) CHECK(scm_ctx_create(0, NULL, &g_ctx));
    return g_ctx;
}

/* scm_use_gpus(n): n > 1 creates one context + NCCL rank per device inside the library (ncclCommInitAll); the host-matrix
 * closures (scm_host) then split the cells over all of them.  n <= 1 goes back to a single device.  Device-resident objects
 * created before the switch belong to the old context: call it first, before anything is uploaded. */
SEXP scm_use_gpus(SEXP n_) {
    int n = Rf_asInteger(n_);
    if (g_multi) { scm_multi_destroy(g_multi); g_multi = NULL; }
    if (n > 1) {
        if (g_ctx) { scm_ctx_destroy(g_ctx); g_ctx = NULL; }
        CHECK(scm_multi_create(NULL, n, &g_multi));
    }
    SEXP out = PROTECT(Rf_allocVector(INTSXP, 1));
    INTEGER(out)[0] = g_multi ? scm_multi_size(g_multi) : 1;
    UNPROTECT(1);
    return out;
}

/* device matrix held by R as an external pointer {ptr, m, n, ld}; freed by the finalizer */
typedef struct { double* p; int64_t m, n, ld; } dmat;
static void dmat_fin(SEXP x) { dmat* d = (dmat*)R_ExternalPtrAddr(x); if (d) { scm_dev_free(ctx(), d->p); R_Free(d); R_ClearExternalPtr(x); } }

/* Y: R numeric matrix.  layout 0: genes x cells (column-major == our row-major cells x genes), 1: cells x genes */
SEXP scm_upload(SEXP Y, SEXP layout) {
    SEXP dim = Rf_getAttrib(Y, R_DimSymbol);
    int64_t r = INTEGER(dim)[0], c = INTEGER(dim)[1];
    dmat* d = R_Calloc(1, dmat);
    int lay = Rf_asInteger(layout);
    d->m = lay == 0 ? c : r; d->n = lay == 0 ? r : c; d->ld = (d->n + 1) / 2 * 2;
    CHECK(scm_dev_alloc(ctx(), d->m * d->ld * 8, (void**)&d->p));
    if (lay == 0) CHECK(scm_copy2d(ctx(), d->p, d->ld * 8, REAL(Y), d->n * 8, d->n * 8, d->m, 0));
    else {
        double* tmp; CHECK(scm_dev_alloc(ctx(), r * c * 8, (void**)&tmp));
        CHECK(scm_h2d(ctx(), tmp, REAL(Y), r * c * 8));
        CHECK(scm_transpose(ctx(), tmp, c, r, r, d->p, d->ld));   /* [gene][cell] -> [cell][gene] */
        CHECK(scm_dev_free(ctx(), tmp));
    }
    CHECK(scm_ctx_sync(ctx()));
    SEXP out = PROTECT(R_MakeExternalPtr(d, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(out, dmat_fin, TRUE);
    UNPROTECT(1);
    return out;
}

static void* to_dev(const void* h, int64_t bytes) { void* d; CHECK(scm_dev_alloc(ctx(), bytes, &d)); CHECK(scm_h2d(ctx(), d, h, bytes)); return d; }

/* plain device buffer held by R (centred Gram, column means); freed by the finalizer */
static void dbuf_fin(SEXP x) { void* p = R_ExternalPtrAddr(x); if (p) { scm_dev_free(ctx(), p); R_ClearExternalPtr(x); } }
static SEXP wrap_dbuf(void* p) {
    SEXP out = PROTECT(R_MakeExternalPtr(p, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(out, dbuf_fin, TRUE);
    UNPROTECT(1);
    return out;
}

/* fit: group (int, 0-based), G, batch (int or NULL), Bt, s, kconv, mode
 *   -> list(fullalpha s x n [row-major], sigma, mean, sd[, gram = <n x n on the device>, colmean = <n on the device>])
 * keep_gram: scm_ruv3_fit_gram keeps the centred Gram of all cells and their column means on the device for scm_sil();
 * more_k (int vector or NULL): further ruvK values the same fullalpha has to serve (scm_ctx_set_conv_ranks, R/scRUVIII.R:64-75). */
static SEXP fit_impl(SEXP Yp, SEXP group, SEXP G, SEXP batch, SEXP Bt, SEXP s_, SEXP kconv, SEXP mode, int keep_gram, SEXP more_k) {
    dmat* d = (dmat*)R_ExternalPtrAddr(Yp);
    int s = Rf_asInteger(s_), has_b = !Rf_isNull(batch);
    int32_t* dg = (int32_t*)to_dev(INTEGER(group), d->m * 4);
    int32_t* db = has_b ? (int32_t*)to_dev(INTEGER(batch), d->m * 4) : NULL;
    double *dfa, *dsig, *dmean = NULL, *dsd = NULL, *dgram = NULL, *dcm = NULL;
    CHECK(scm_dev_alloc(ctx(), (int64_t)s * d->n * 8, (void**)&dfa)); CHECK(scm_dev_alloc(ctx(), s * 8, (void**)&dsig));
    if (has_b) { CHECK(scm_dev_alloc(ctx(), d->n * 8, (void**)&dmean)); CHECK(scm_dev_alloc(ctx(), d->n * 8, (void**)&dsd)); }
    if (keep_gram) { CHECK(scm_dev_alloc(ctx(), d->n * d->n * 8, (void**)&dgram)); CHECK(scm_dev_alloc(ctx(), d->n * 8, (void**)&dcm)); }
    if (!Rf_isNull(more_k)) CHECK(scm_ctx_set_conv_ranks(ctx(), INTEGER(more_k), Rf_length(more_k)));
    int32_t iters; double resid;
    int rc = scm_ruv3_fit_gram(ctx(), d->p, d->m, d->n, d->ld, dg, Rf_asInteger(G), db, has_b ? Rf_asInteger(Bt) : 0, s, Rf_asInteger(kconv),
                               Rf_asInteger(mode), 1e-10, 300, 1, dfa, dsig, dmean, dsd, &iters, &resid, dgram, dcm);
    scm_ctx_set_conv_ranks(ctx(), NULL, 0);
    CHECK(rc);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, keep_gram ? 6 : 4));
    SEXP fa = PROTECT(Rf_allocMatrix(REALSXP, (int)d->n, s));          /* n x s column-major == s x n row-major; R side t()s it */
    CHECK(scm_d2h(ctx(), REAL(fa), dfa, (int64_t)s * d->n * 8)); SET_VECTOR_ELT(out, 0, fa);
    SEXP sg = PROTECT(Rf_allocVector(REALSXP, s)); CHECK(scm_d2h(ctx(), REAL(sg), dsig, s * 8)); SET_VECTOR_ELT(out, 1, sg);
    if (has_b) {
        SEXP mu = PROTECT(Rf_allocVector(REALSXP, d->n)); CHECK(scm_d2h(ctx(), REAL(mu), dmean, d->n * 8)); SET_VECTOR_ELT(out, 2, mu);
        SEXP sd = PROTECT(Rf_allocVector(REALSXP, d->n)); CHECK(scm_d2h(ctx(), REAL(sd), dsd, d->n * 8)); SET_VECTOR_ELT(out, 3, sd);
        UNPROTECT(2);
    }
    if (keep_gram) {
        SEXP nm = PROTECT(Rf_allocVector(STRSXP, 6));
        const char* names[6] = {"fullalpha", "sigma", "mean", "sd", "gram", "colmean"};
        for (int i = 0; i < 6; i++) SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
        Rf_setAttrib(out, R_NamesSymbol, nm);
        SET_VECTOR_ELT(out, 4, wrap_dbuf(dgram)); SET_VECTOR_ELT(out, 5, wrap_dbuf(dcm));
        UNPROTECT(1);
    }
    scm_dev_free(ctx(), dg); scm_dev_free(ctx(), db); scm_dev_free(ctx(), dfa); scm_dev_free(ctx(), dsig);
    scm_dev_free(ctx(), dmean); scm_dev_free(ctx(), dsd);
    UNPROTECT(3);
    return out;
}
SEXP scm_fit(SEXP Yp, SEXP group, SEXP G, SEXP batch, SEXP Bt, SEXP s_, SEXP kconv, SEXP mode) {
    return fit_impl(Yp, group, G, batch, Bt, s_, kconv, mode, 0, R_NilValue);
}
SEXP scm_fit_gram(SEXP Yp, SEXP group, SEXP G, SEXP batch, SEXP Bt, SEXP s_, SEXP kconv, SEXP mode, SEXP more_k) {
    return fit_impl(Yp, group, G, batch, Bt, s_, kconv, mode, 1, more_k);
}

/* hands the library's cached scratch back to the driver (scm_ctx_release_scratch), e.g. from gc() hooks */
SEXP scm_release_scratch(void) { CHECK(scm_ctx_release_scratch(ctx())); return R_NilValue; }

/* adjust: alpha (n x k column-major == k x n row-major), ctl (logical n), center/pre/post (numeric n or NULL), subset (int 0-based or NULL)
 * returns an R matrix genes(or subset) x cells, i.e. the byte image of the device result */
SEXP scm_adjust_(SEXP Yp, SEXP alpha, SEXP k_, SEXP ctl, SEXP center, SEXP pre, SEXP post, SEXP subset) {
    dmat* d = (dmat*)R_ExternalPtrAddr(Yp);
    int k = Rf_asInteger(k_), nsub = Rf_isNull(subset) ? 0 : Rf_length(subset);
    double* da = (double*)to_dev(REAL(alpha), (int64_t)k * d->n * 8);
    uint8_t* hc = (uint8_t*)R_alloc(d->n, 1); for (int64_t j = 0; j < d->n; j++) hc[j] = LOGICAL(ctl)[j] != 0;
    uint8_t* dc = (uint8_t*)to_dev(hc, d->n);
    double* dce = Rf_isNull(center) ? NULL : (double*)to_dev(REAL(center), d->n * 8);
    double* dpr = Rf_isNull(pre) ? NULL : (double*)to_dev(REAL(pre), d->n * 8);
    double* dpo = Rf_isNull(post) ? NULL : (double*)to_dev(REAL(post), d->n * 8);
    int32_t* dsub = nsub ? (int32_t*)to_dev(INTEGER(subset), nsub * 4) : NULL;
    scm_coef* coef; CHECK(scm_coef_create(ctx(), da, d->n, k, dc, dce, dpr, dpo, &coef));
    int64_t nout = nsub ? nsub : d->n, ldo = (nout + 1) / 2 * 2;
    double* dout; CHECK(scm_dev_alloc(ctx(), d->m * ldo * 8, (void**)&dout));
    CHECK(scm_adjust(ctx(), d->p, d->m, d->n, d->ld, coef, dsub, nsub, dout, ldo, NULL));
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)nout, (int)d->m));
    CHECK(scm_copy2d(ctx(), REAL(out), nout * 8, dout, ldo * 8, nout * 8, d->m, 1));
    scm_coef_destroy(ctx(), coef);
    scm_dev_free(ctx(), da); scm_dev_free(ctx(), dc); scm_dev_free(ctx(), dce); scm_dev_free(ctx(), dpr); scm_dev_free(ctx(), dpo);
    scm_dev_free(ctx(), dsub); scm_dev_free(ctx(), dout);
    UNPROTECT(1);
    return out;
}

/* segmented means by key: returns G x n (as n x G column-major) and counts */
SEXP scm_segmean(SEXP Yp, SEXP key, SEXP G_) {
    dmat* d = (dmat*)R_ExternalPtrAddr(Yp);
    int G = Rf_asInteger(G_);
    int32_t* dk = (int32_t*)to_dev(INTEGER(key), d->m * 4);
    double* dm; int64_t* dc;
    CHECK(scm_dev_alloc(ctx(), (int64_t)G * d->n * 8, (void**)&dm)); CHECK(scm_dev_alloc(ctx(), G * 8, (void**)&dc));
    CHECK(scm_seg_colmean(ctx(), d->p, d->m, d->n, d->ld, dk, G, dm, dc, NULL));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP mm = PROTECT(Rf_allocMatrix(REALSXP, (int)d->n, G)); CHECK(scm_d2h(ctx(), REAL(mm), dm, (int64_t)G * d->n * 8));
    int64_t* hc = (int64_t*)R_alloc(G, 8); CHECK(scm_d2h(ctx(), hc, dc, G * 8));
    SEXP cc = PROTECT(Rf_allocVector(REALSXP, G)); for (int g = 0; g < G; g++) REAL(cc)[g] = (double)hc[g];
    SET_VECTOR_ELT(out, 0, mm); SET_VECTOR_ELT(out, 1, cc);
    scm_dev_free(ctx(), dk); scm_dev_free(ctx(), dm); scm_dev_free(ctx(), dc);
    UNPROTECT(3);
    return out;
}

/* dgCMatrix (genes x cells: slots @p int, @i int, @x double, @Dim) -> dense, cosine-normalised cell rows on the device (K8).
 * Replaces methods::as(., "CsparseMatrix") + batchelor::cosineNorm (R/scMerge2.R:164-166,193-198) + the dense upload. */
SEXP scm_upload_sparse(SEXP p, SEXP i, SEXP x, SEXP dim, SEXP cosine) {
    int64_t n = INTEGER(dim)[0], m = INTEGER(dim)[1], nnz = Rf_xlength(x);
    int64_t* hp = (int64_t*)R_alloc(m + 1, 8);
    for (int64_t c = 0; c <= m; c++) hp[c] = INTEGER(p)[c];               /* @p is 32-bit in R: widen */
    int64_t* dp = (int64_t*)to_dev(hp, (m + 1) * 8);
    int32_t* di = (int32_t*)to_dev(INTEGER(i), nnz * 4);
    double* dx = (double*)to_dev(REAL(x), nnz * 8);
    int32_t hflag = 0, *dflag = (int32_t*)to_dev(&hflag, 4);
    dmat* d = R_Calloc(1, dmat);
    d->m = m; d->n = n; d->ld = (n + 1) / 2 * 2;
    CHECK(scm_dev_alloc(ctx(), d->m * d->ld * 8, (void**)&d->p));
    CHECK(scm_csc_to_dense(ctx(), dp, di, dx, m, n, nnz, Rf_asLogical(cosine), 1e-8, d->p, d->ld, NULL, dflag));
    CHECK(scm_d2h(ctx(), &hflag, dflag, 4));
    scm_dev_free(ctx(), dp); scm_dev_free(ctx(), di); scm_dev_free(ctx(), dx); scm_dev_free(ctx(), dflag);
    if (hflag) Rf_warning("Y contains missing values or infinities.");       /* pseudoRUVIII only warns, R/scMerge2_utilis.R:49-52 */
    SEXP out = PROTECT(R_MakeExternalPtr(d, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(out, dmat_fin, TRUE);
    UNPROTECT(1);
    return out;
}

/* scRUVg (R/scRUVg.R:30-85, Z = 1, eta = NULL): list(newY genes x cells image, W (k x m image), alpha (n x k image)) */
SEXP scm_ruvg(SEXP Yp, SEXP ctl, SEXP k_) {
    dmat* d = (dmat*)R_ExternalPtrAddr(Yp);
    int k = Rf_asInteger(k_);
    uint8_t* hc = (uint8_t*)R_alloc(d->n, 1); for (int64_t j = 0; j < d->n; j++) hc[j] = LOGICAL(ctl)[j] != 0;
    uint8_t* dc = (uint8_t*)to_dev(hc, d->n);
    double *dal, *dce, *dout, *dW; scm_coef* coef; int32_t iters; double resid;
    CHECK(scm_dev_alloc(ctx(), (int64_t)k * d->n * 8, (void**)&dal)); CHECK(scm_dev_alloc(ctx(), d->n * 8, (void**)&dce));
    CHECK(scm_dev_alloc(ctx(), d->m * d->ld * 8, (void**)&dout)); CHECK(scm_dev_alloc(ctx(), d->m * k * 8, (void**)&dW));
    CHECK(scm_ruvg_fit(ctx(), d->p, d->m, d->n, d->ld, dc, k, SCM_MODE_EXACT, 1e-10, 500, 1, dal, dce, NULL, &coef, &iters, &resid));
    CHECK(scm_adjust(ctx(), d->p, d->m, d->n, d->ld, coef, NULL, 0, dout, d->ld, dW));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP ny = PROTECT(Rf_allocMatrix(REALSXP, (int)d->n, (int)d->m));
    CHECK(scm_copy2d(ctx(), REAL(ny), d->n * 8, dout, d->ld * 8, d->n * 8, d->m, 1)); SET_VECTOR_ELT(out, 0, ny);
    SEXP W = PROTECT(Rf_allocMatrix(REALSXP, k, (int)d->m)); CHECK(scm_d2h(ctx(), REAL(W), dW, d->m * k * 8)); SET_VECTOR_ELT(out, 1, W);
    SEXP al = PROTECT(Rf_allocMatrix(REALSXP, (int)d->n, k)); CHECK(scm_d2h(ctx(), REAL(al), dal, (int64_t)k * d->n * 8)); SET_VECTOR_ELT(out, 2, al);
    scm_coef_destroy(ctx(), coef);
    scm_dev_free(ctx(), dc); scm_dev_free(ctx(), dal); scm_dev_free(ctx(), dce); scm_dev_free(ctx(), dout); scm_dev_free(ctx(), dW);
    UNPROTECT(4);
    return out;
}

/* ruv::RUV1 on a device matrix, in place: Y <- Y - Y[, ctl] etac' (etac etac')^-1 eta (R/fastRUVIII.R:63-64, R/scMerge2_utilis.R:56).
 * eta_t: n x q image of the q x n design rows.  `copy` != 0: the result goes to a NEW device matrix (the handle given stays as it is). */
SEXP scm_ruv1_(SEXP Yp, SEXP eta_t, SEXP q_, SEXP ctl, SEXP copy) {
    dmat* d = (dmat*)R_ExternalPtrAddr(Yp);
    int q = Rf_asInteger(q_);
    double* de = (double*)to_dev(REAL(eta_t), (int64_t)q * d->n * 8);
    uint8_t* hc = (uint8_t*)R_alloc(d->n, 1); for (int64_t j = 0; j < d->n; j++) hc[j] = LOGICAL(ctl)[j] != 0;
    uint8_t* dc = (uint8_t*)to_dev(hc, d->n);
    scm_coef* coef; CHECK(scm_coef_create(ctx(), de, d->n, q, dc, NULL, NULL, NULL, &coef));
    SEXP out = Yp;
    if (Rf_asInteger(copy)) {
        dmat* o = R_Calloc(1, dmat); *o = *d;
        CHECK(scm_dev_alloc(ctx(), d->m * d->ld * 8, (void**)&o->p));
        CHECK(scm_adjust(ctx(), d->p, d->m, d->n, d->ld, coef, NULL, 0, o->p, o->ld, NULL));
        out = PROTECT(R_MakeExternalPtr(o, R_NilValue, R_NilValue));
        R_RegisterCFinalizerEx(out, dmat_fin, TRUE);
        UNPROTECT(1);
    } else {
        CHECK(scm_adjust(ctx(), d->p, d->m, d->n, d->ld, coef, NULL, 0, d->p, d->ld, NULL));
    }
    scm_coef_destroy(ctx(), coef); scm_dev_free(ctx(), de); scm_dev_free(ctx(), dc);
    return out;
}

/* scRUVg with eta and / or a re-used fullW (R/scRUVg.R:59,64,77-80).  Yp: the original matrix, Y0p: RUV1(Y) (the same handle when
 * eta is NULL), Wt: k x m image of W = fullW[, 1:k] or NULL (then W comes from the decomposition of Y0). */
SEXP scm_ruvg2(SEXP Yp, SEXP Y0p, SEXP ctl, SEXP k_, SEXP Wt) {
    dmat* d = (dmat*)R_ExternalPtrAddr(Yp); dmat* d0 = (dmat*)R_ExternalPtrAddr(Y0p);
    int k = Rf_asInteger(k_);
    uint8_t* hc = (uint8_t*)R_alloc(d->n, 1); for (int64_t j = 0; j < d->n; j++) hc[j] = LOGICAL(ctl)[j] != 0;
    uint8_t* dc = (uint8_t*)to_dev(hc, d->n);
    double *dal, *dce, *dout, *dW; scm_coef* coef = NULL; int32_t iters; double resid;
    CHECK(scm_dev_alloc(ctx(), (int64_t)k * d->n * 8, (void**)&dal)); CHECK(scm_dev_alloc(ctx(), d->n * 8, (void**)&dce));
    CHECK(scm_dev_alloc(ctx(), d->m * d->ld * 8, (void**)&dout)); CHECK(scm_dev_alloc(ctx(), d->m * k * 8, (void**)&dW));
    if (Rf_isNull(Wt)) {
        CHECK(scm_ruvg_fit(ctx(), d0->p, d0->m, d0->n, d0->ld, dc, k, SCM_MODE_EXACT, 1e-10, 500, 1, dal, dce, NULL, &coef, &iters, &resid));
        CHECK(scm_adjust(ctx(), d0->p, d0->m, d0->n, d0->ld, coef, NULL, 0, NULL, 0, dW));     /* projection only: W = (Y0 - mean) B */
    } else {
        CHECK(scm_h2d(ctx(), dW, REAL(Wt), d->m * k * 8));
        int32_t* dk; int64_t* dcnt;
        CHECK(scm_dev_alloc(ctx(), d->m * 4, (void**)&dk)); CHECK(scm_dev_alloc(ctx(), 8, (void**)&dcnt));
        int32_t* hk = (int32_t*)R_alloc(d->m, 4); for (int64_t i = 0; i < d->m; i++) hk[i] = 0;
        CHECK(scm_h2d(ctx(), dk, hk, d->m * 4));
        CHECK(scm_seg_colmean(ctx(), d0->p, d0->m, d0->n, d0->ld, dk, 1, dce, dcnt, NULL));   /* colMeans(Y0): Z = 1 */
        CHECK(scm_ls_coef(ctx(), d0->p, d0->m, d0->n, d0->ld, dW, k, dce, dal));
        scm_dev_free(ctx(), dk); scm_dev_free(ctx(), dcnt);
    }
    CHECK(scm_rankk_subtract(ctx(), d->p, d->m, d->n, d->ld, dW, k, dal, dout, d->ld));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP ny = PROTECT(Rf_allocMatrix(REALSXP, (int)d->n, (int)d->m));
    CHECK(scm_copy2d(ctx(), REAL(ny), d->n * 8, dout, d->ld * 8, d->n * 8, d->m, 1)); SET_VECTOR_ELT(out, 0, ny);
    SEXP W = PROTECT(Rf_allocMatrix(REALSXP, k, (int)d->m)); CHECK(scm_d2h(ctx(), REAL(W), dW, d->m * k * 8)); SET_VECTOR_ELT(out, 1, W);
    SEXP al = PROTECT(Rf_allocMatrix(REALSXP, (int)d->n, k)); CHECK(scm_d2h(ctx(), REAL(al), dal, (int64_t)k * d->n * 8)); SET_VECTOR_ELT(out, 2, al);
    if (coef) scm_coef_destroy(ctx(), coef);
    scm_dev_free(ctx(), dc); scm_dev_free(ctx(), dal); scm_dev_free(ctx(), dce); scm_dev_free(ctx(), dout); scm_dev_free(ctx(), dW);
    UNPROTECT(4);
    return out;
}

/* choice of ruvK (R/scRUVIII.R:83-105): for the candidate described by (alpha, k, ctl, center, pre, post) returns
 * c(silhouette(cell_type), silhouette(batch)) in the first `rank` PCs of the adjusted matrix.  gram / colmean are the external
 * pointers returned by scm_fit_gram (scm_ruv3_fit_gram); labels are 0-based ints. */
SEXP scm_sil(SEXP Yp, SEXP gram, SEXP colmean, SEXP alpha, SEXP k_, SEXP ctl, SEXP center, SEXP pre, SEXP post, SEXP ct, SEXP nct,
             SEXP bt, SEXP nbt, SEXP rank_) {
    dmat* d = (dmat*)R_ExternalPtrAddr(Yp);
    int k = Rf_asInteger(k_), rank = Rf_asInteger(rank_);
    double* da = (double*)to_dev(REAL(alpha), (int64_t)k * d->n * 8);
    uint8_t* hc = (uint8_t*)R_alloc(d->n, 1); for (int64_t j = 0; j < d->n; j++) hc[j] = LOGICAL(ctl)[j] != 0;
    uint8_t* dc = (uint8_t*)to_dev(hc, d->n);
    double* dce = Rf_isNull(center) ? NULL : (double*)to_dev(REAL(center), d->n * 8);
    double* dpr = Rf_isNull(pre) ? NULL : (double*)to_dev(REAL(pre), d->n * 8);
    double* dpo = Rf_isNull(post) ? NULL : (double*)to_dev(REAL(post), d->n * 8);
    int32_t* dct = (int32_t*)to_dev(INTEGER(ct), d->m * 4);
    int32_t* dbt = (int32_t*)to_dev(INTEGER(bt), d->m * 4);
    scm_coef* coef; CHECK(scm_coef_create(ctx(), da, d->n, k, dc, dce, dpr, dpo, &coef));
    double* dsc; CHECK(scm_dev_alloc(ctx(), d->m * rank * 8, (void**)&dsc));
    int32_t iters; double resid, sa, sb;
    CHECK(scm_pca_scores(ctx(), (const double*)R_ExternalPtrAddr(gram), d->n, d->m, coef, (const double*)R_ExternalPtrAddr(colmean), rank,
                         d->p, d->m, d->ld, dsc, NULL, 1e-9, 500, &iters, &resid));
    CHECK(scm_silhouette(ctx(), dsc, d->m, rank, rank, dct, Rf_asInteger(nct), dbt, Rf_asInteger(nbt), 0, d->m, &sa, &sb));
    scm_coef_destroy(ctx(), coef);
    scm_dev_free(ctx(), da); scm_dev_free(ctx(), dc); scm_dev_free(ctx(), dce); scm_dev_free(ctx(), dpr); scm_dev_free(ctx(), dpo);
    scm_dev_free(ctx(), dct); scm_dev_free(ctx(), dbt); scm_dev_free(ctx(), dsc);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 2));
    REAL(out)[0] = sa / (double)d->m; REAL(out)[1] = sb / (double)d->m;
    UNPROTECT(1);
    return out;
}

/* The whole closure on the R matrix itself, PCIe overlapped with the kernels (scm_ruv3_host / scm_multi_ruv3_host): upload in
 * chunks with the fit streaming behind, adjust + download chunk by chunk.  REAL(Y) is ordinary pageable memory; the library
 * stages it through its own page-locked bounce buffers with helper threads, so no pinning is needed on the R side.
 *   Y      : genes x cells numeric matrix (column-major == row-major cells x genes; the layout scMerge()/scMerge2() hold)
 *   group  : int 0-based replicate keys or NULL (adjust only);  batch: int keys or NULL (non-NULL: scRUVIII)
 *   alpha_in: NULL = fit, else s_in x n fullalpha (as n x s_in column-major) for adjust-only;  center_mode 0 none, 1 colMeans, 2 center
 * returns list(newY genes x cells, fullalpha (n x s image) | NULL, sigma | NULL, mean | NULL, sd | NULL, conv_rank) */
SEXP scm_host(SEXP Y, SEXP group, SEXP G, SEXP batch, SEXP Bt, SEXP ctl, SEXP k_, SEXP s_, SEXP mode, SEXP alpha_in, SEXP center_mode,
              SEXP center) {
    SEXP dim = Rf_getAttrib(Y, R_DimSymbol);
    int64_t n = INTEGER(dim)[0], m = INTEGER(dim)[1];
    int k = Rf_asInteger(k_), fit = Rf_isNull(alpha_in), has_b = !Rf_isNull(batch);
    int s = fit ? Rf_asInteger(s_) : (int)(Rf_xlength(alpha_in) / n);
    uint8_t* hc = (uint8_t*)R_alloc(n, 1); for (int64_t j = 0; j < n; j++) hc[j] = LOGICAL(ctl)[j] != 0;
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 6));
    SEXP newY = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)m)); SET_VECTOR_ELT(out, 0, newY);
    SEXP fa = R_NilValue, sg = R_NilValue, mu = R_NilValue, sd = R_NilValue;
    int np = 2;
    if (fit) {
        fa = PROTECT(Rf_allocMatrix(REALSXP, (int)n, s)); SET_VECTOR_ELT(out, 1, fa);
        sg = PROTECT(Rf_allocVector(REALSXP, s)); SET_VECTOR_ELT(out, 2, sg); np += 2;
    }
    if (has_b || Rf_asInteger(center_mode) == 1) { mu = PROTECT(Rf_allocVector(REALSXP, n)); SET_VECTOR_ELT(out, 3, mu); np++; }
    if (has_b) { sd = PROTECT(Rf_allocVector(REALSXP, n)); SET_VECTOR_ELT(out, 4, sd); np++; }
    int32_t iters = 0, conv = 0; double resid = 0.0;
    const int32_t* hg = Rf_isNull(group) ? NULL : (const int32_t*)INTEGER(group);
    const int32_t* hb = has_b ? (const int32_t*)INTEGER(batch) : NULL;
    const double* hin = fit ? NULL : REAL(alpha_in);
    const double* hcen = Rf_isNull(center) ? NULL : REAL(center);
    int rc;
    if (g_multi)
        rc = scm_multi_ruv3_host(g_multi, REAL(Y), m, n, n, hg, Rf_asInteger(G), hb, has_b ? Rf_asInteger(Bt) : 0, hc, k, s, Rf_asInteger(mode),
                                 1e-10, 500, 1, hin, fit ? 0 : s, Rf_asInteger(center_mode), hcen, REAL(newY), n,
                                 fit ? REAL(fa) : NULL, fit ? REAL(sg) : NULL, mu == R_NilValue ? NULL : REAL(mu),
                                 sd == R_NilValue ? NULL : REAL(sd), &iters, &resid);
    else
        rc = scm_ruv3_host(ctx(), REAL(Y), m, n, n, hg, Rf_asInteger(G), hb, has_b ? Rf_asInteger(Bt) : 0, hc, k, s, Rf_asInteger(mode),
                           1e-10, 500, 1, hin, fit ? 0 : s, Rf_asInteger(center_mode), hcen, REAL(newY), n,
                           fit ? REAL(fa) : NULL, fit ? REAL(sg) : NULL, mu == R_NilValue ? NULL : REAL(mu),
                           sd == R_NilValue ? NULL : REAL(sd), &iters, &resid);
    CHECK(rc);
    CHECK(scm_ctx_fit_info(ctx(), NULL, NULL, &conv, NULL));
    SEXP cr = PROTECT(Rf_allocVector(INTSXP, 1)); INTEGER(cr)[0] = fit ? conv : NA_INTEGER; SET_VECTOR_ELT(out, 5, cr); np++;
    UNPROTECT(np);
    return out;
}

/* ---- sparse-native scMerge2 passes: the dgCMatrix slots stay the only copy of the cells on the device ---------------------
 * scm_upload_csc: (@p, @i, @x, @Dim of a genes x cells dgCMatrix, cosineNorm) -> external pointer {indptr, idx, val, inv_norm}
 * (validated with scm_csc_check; un-canonical input is an error here: use scm_upload_sparse for the dense path). */
typedef struct { int64_t* p; int32_t* i; double* x; double* inv; int64_t m, n, nnz; } dcsc;
static void dcsc_fin(SEXP e) {
    dcsc* d = (dcsc*)R_ExternalPtrAddr(e);
    if (d) { scm_dev_free(ctx(), d->p); scm_dev_free(ctx(), d->i); scm_dev_free(ctx(), d->x); scm_dev_free(ctx(), d->inv); R_Free(d); R_ClearExternalPtr(e); }
}
SEXP scm_upload_csc(SEXP p, SEXP i, SEXP x, SEXP dim, SEXP cosine) {
    int64_t n = INTEGER(dim)[0], m = INTEGER(dim)[1], nnz = Rf_xlength(x);
    int64_t* hp = (int64_t*)R_alloc(m + 1, 8);
    for (int64_t c = 0; c <= m; c++) hp[c] = INTEGER(p)[c];
    dcsc* d = R_Calloc(1, dcsc);
    d->m = m; d->n = n; d->nnz = nnz;
    d->p = (int64_t*)to_dev(hp, (m + 1) * 8); d->i = (int32_t*)to_dev(INTEGER(i), nnz * 4); d->x = (double*)to_dev(REAL(x), nnz * 8);
    int32_t canonical = 0, hflag = 0, *dflag = (int32_t*)to_dev(&hflag, 4);
    CHECK(scm_csc_check(ctx(), d->p, d->i, m, n, nnz, &canonical));
    if (!canonical) Rf_error("cells with unsorted or duplicate gene indices: not a canonical dgCMatrix");
    CHECK(scm_dev_alloc(ctx(), m * 8, (void**)&d->inv));
    CHECK(scm_csc_rownorm(ctx(), d->p, d->x, m, Rf_asLogical(cosine), 1e-8, d->inv, dflag));
    CHECK(scm_d2h(ctx(), &hflag, dflag, 4));
    scm_dev_free(ctx(), dflag);
    if (hflag) Rf_warning("Y contains missing values or infinities.");
    SEXP out = PROTECT(R_MakeExternalPtr(d, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(out, dcsc_fin, TRUE);
    UNPROTECT(1);
    return out;
}
/* pseudobulk means by int key (create_pseudoBulk, R/scMerge2_utilis.R:461-486) + colMeans of all cells (:97):
 * list(means n x P image, counts P, colmeans n) */
SEXP scm_csc_pseudobulk(SEXP cp, SEXP key, SEXP P_) {
    dcsc* d = (dcsc*)R_ExternalPtrAddr(cp);
    int P = Rf_asInteger(P_);
    int64_t ldn = (d->n + 1) / 2 * 2;
    int32_t* dk = (int32_t*)to_dev(INTEGER(key), d->m * 4);
    double *dm, *dmean; int64_t* dc;
    CHECK(scm_dev_alloc(ctx(), (int64_t)P * ldn * 8, (void**)&dm)); CHECK(scm_dev_alloc(ctx(), P * 8, (void**)&dc));
    CHECK(scm_dev_alloc(ctx(), d->n * 8, (void**)&dmean));
    CHECK(scm_csc_seg_colmean(ctx(), d->p, d->i, d->x, d->inv, d->m, d->n, dk, P, dm, ldn, dc));
    CHECK(scm_group_colmean(ctx(), dm, dc, P, d->n, ldn, dmean));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP mm = PROTECT(Rf_allocMatrix(REALSXP, (int)d->n, P));
    CHECK(scm_copy2d(ctx(), REAL(mm), d->n * 8, dm, ldn * 8, d->n * 8, P, 1));
    int64_t* hcn = (int64_t*)R_alloc(P, 8); CHECK(scm_d2h(ctx(), hcn, dc, (int64_t)P * 8));
    SEXP cc = PROTECT(Rf_allocVector(REALSXP, P)); for (int g = 0; g < P; g++) REAL(cc)[g] = (double)hcn[g];
    SEXP cm = PROTECT(Rf_allocVector(REALSXP, d->n)); CHECK(scm_d2h(ctx(), REAL(cm), dmean, d->n * 8));
    SET_VECTOR_ELT(out, 0, mm); SET_VECTOR_ELT(out, 1, cc); SET_VECTOR_ELT(out, 2, cm);
    scm_dev_free(ctx(), dk); scm_dev_free(ctx(), dm); scm_dev_free(ctx(), dc); scm_dev_free(ctx(), dmean);
    UNPROTECT(4);
    return out;
}
/* newY (genes x cells) = adjust of every cell from the sparse rows, result streamed to the R matrix chunk by chunk
 * (scm_csc_adjust_to_host: the dense matrix never exists on the device).  alpha: n x k image, ctl logical n, center numeric n. */
SEXP scm_csc_adjust_(SEXP cp, SEXP alpha, SEXP k_, SEXP ctl, SEXP center) {
    dcsc* d = (dcsc*)R_ExternalPtrAddr(cp);
    int k = Rf_asInteger(k_);
    double* da = (double*)to_dev(REAL(alpha), (int64_t)k * d->n * 8);
    uint8_t* hc = (uint8_t*)R_alloc(d->n, 1); for (int64_t j = 0; j < d->n; j++) hc[j] = LOGICAL(ctl)[j] != 0;
    uint8_t* dc = (uint8_t*)to_dev(hc, d->n);
    double* dce = Rf_isNull(center) ? NULL : (double*)to_dev(REAL(center), d->n * 8);
    scm_coef* coef; CHECK(scm_coef_create(ctx(), da, d->n, k, dc, dce, NULL, NULL, &coef));
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)d->n, (int)d->m));
    CHECK(scm_csc_adjust_to_host(ctx(), d->p, d->i, d->x, d->inv, d->m, d->n, coef, REAL(out), d->n));
    scm_coef_destroy(ctx(), coef);
    scm_dev_free(ctx(), da); scm_dev_free(ctx(), dc); scm_dev_free(ctx(), dce);
    UNPROTECT(1);
    return out;
}

/* ---- replicate-finding numerics of scReplicate (R/scReplicate.R:458-520,793-819; include/scmerge_b200.h "replicate-finding") -----
 * X arguments are genes x cells R matrices (column-major == the device's row-major cells x genes). */
static dmat up_gc(SEXP X) {   /* temporary upload of a genes x cells matrix; caller frees d.p */
    SEXP dim = Rf_getAttrib(X, R_DimSymbol);
    dmat d; d.n = INTEGER(dim)[0]; d.m = INTEGER(dim)[1]; d.ld = (d.n + 1) / 2 * 2;
    CHECK(scm_dev_alloc(ctx(), d.m * d.ld * 8, (void**)&d.p));
    CHECK(scm_copy2d(ctx(), d.p, d.ld * 8, REAL(X), d.n * 8, d.n * 8, d.m, 0));
    return d;
}
/* squared distance of every cell to the per-gene MEDIAN of its cluster (centroidDist before rank(), R/scReplicate.R:459-460) */
SEXP scm_centroid_sqdist(SEXP X, SEXP key, SEXP G_) {
    dmat d = up_gc(X);
    int G = Rf_asInteger(G_);
    int32_t* dk = (int32_t*)to_dev(INTEGER(key), d.m * 4);
    double *dmed, *ddist;
    CHECK(scm_dev_alloc(ctx(), (int64_t)G * d.n * 8, (void**)&dmed)); CHECK(scm_dev_alloc(ctx(), d.m * 8, (void**)&ddist));
    CHECK(scm_group_median(ctx(), d.p, d.m, d.n, d.ld, (const int32_t*)INTEGER(key), dk, G, dmed));
    CHECK(scm_sqdist_to_center(ctx(), d.p, d.m, d.n, d.ld, dk, G, dmed, ddist));
    SEXP out = PROTECT(Rf_allocVector(REALSXP, d.m));
    CHECK(scm_d2h(ctx(), REAL(out), ddist, d.m * 8));
    scm_dev_free(ctx(), dk); scm_dev_free(ctx(), dmed); scm_dev_free(ctx(), ddist); scm_dev_free(ctx(), d.p);
    UNPROTECT(1);
    return out;
}
/* compute_dist_mat (R/scReplicate.R:470-490): returns the m2 x m1 image of the m1 x m2 row-major distance matrix (R side t()s it);
 * with offsets (numeric vectors K1 + 1 / K2 + 1, cells ordered by cluster) the K2 x K1 image of the block medians of compute_dist_res (:507-513) */
SEXP scm_dist(SEXP X1, SEXP X2, SEXP metric, SEXP row_off, SEXP col_off) {
    dmat a = up_gc(X1), b = up_gc(X2);
    double* dD; CHECK(scm_dev_alloc(ctx(), a.m * b.m * 8, (void**)&dD));
    CHECK(scm_pair_dist(ctx(), a.p, a.m, a.ld, b.p, b.m, b.ld, a.n, Rf_asInteger(metric), dD, b.m));
    SEXP out;
    if (Rf_isNull(row_off)) {
        out = PROTECT(Rf_allocMatrix(REALSXP, (int)b.m, (int)a.m));
        CHECK(scm_d2h(ctx(), REAL(out), dD, a.m * b.m * 8));
    } else {
        int K1 = Rf_length(row_off) - 1, K2 = Rf_length(col_off) - 1;
        int64_t* ro = (int64_t*)R_alloc(K1 + 1, 8); int64_t* co = (int64_t*)R_alloc(K2 + 1, 8);
        for (int i = 0; i <= K1; i++) ro[i] = (int64_t)REAL(row_off)[i];
        for (int j = 0; j <= K2; j++) co[j] = (int64_t)REAL(col_off)[j];
        out = PROTECT(Rf_allocMatrix(REALSXP, K2, K1));
        CHECK(scm_block_median(ctx(), dD, b.m, ro, K1, co, K2, REAL(out)));
    }
    scm_dev_free(ctx(), dD); scm_dev_free(ctx(), a.p); scm_dev_free(ctx(), b.p);
    UNPROTECT(1);
    return out;
}
/* BiocSingular::runPCA(t(X), rank, scale = TRUE, center = TRUE)$x (R/scReplicate.R:813-815): rank x cells image of the score matrix */
SEXP scm_pca(SEXP X, SEXP rank_) {
    dmat d = up_gc(X);
    int rank = Rf_asInteger(rank_);
    double *dg, *dm_, *dsc; int64_t mt = 0; int32_t iters; double resid;
    CHECK(scm_dev_alloc(ctx(), d.n * d.n * 8, (void**)&dg)); CHECK(scm_dev_alloc(ctx(), d.n * 8, (void**)&dm_));
    CHECK(scm_dev_alloc(ctx(), d.m * rank * 8, (void**)&dsc));
    CHECK(scm_centred_gram(ctx(), d.p, d.m, d.n, d.ld, dg, dm_, &mt));
    CHECK(scm_pca_scores(ctx(), dg, d.n, mt, NULL, dm_, rank, d.p, d.m, d.ld, dsc, NULL, 1e-9, 500, &iters, &resid));
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, rank, (int)d.m));
    CHECK(scm_d2h(ctx(), REAL(out), dsc, d.m * rank * 8));
    scm_dev_free(ctx(), dg); scm_dev_free(ctx(), dm_); scm_dev_free(ctx(), dsc); scm_dev_free(ctx(), d.p);
    UNPROTECT(1);
    return out;
}

/* stats::kmeans(t(Xt), centers = K, nstart = nstart) of identifyCluster (R/scReplicate.R:394-395) on the device.  Xt: coordinates x
 * points (the transposed score matrix: R's column-major memory is then points x coordinates by rows).  Returns
 * list(cluster (1-based), centers (K x d), withinss, tot.withinss, size, iter, ifault). */
SEXP scm_kmeans_(SEXP Xt, SEXP K_, SEXP nstart_, SEXP iter_max_, SEXP seed_) {
    dmat d = up_gc(Xt);
    int K = Rf_asInteger(K_), iter_max = Rf_asInteger(iter_max_);
    int32_t* dl; double* dc;
    CHECK(scm_dev_alloc(ctx(), d.m * 4, (void**)&dl)); CHECK(scm_dev_alloc(ctx(), (int64_t)K * d.n * 8, (void**)&dc));
    double* wss = (double*)R_alloc(K, 8); int64_t* size = (int64_t*)R_alloc(K, 8);
    double tot = 0.0; int32_t info[4];
    CHECK(scm_kmeans(ctx(), d.p, d.m, (int32_t)d.n, d.ld, K, Rf_asInteger(nstart_), iter_max, (uint64_t)Rf_asReal(seed_), dl, dc, wss, size, &tot, info));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 7));
    SEXP cl = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)d.m));
    CHECK(scm_d2h(ctx(), INTEGER(cl), dl, d.m * 4));
    for (int64_t i = 0; i < d.m; i++) INTEGER(cl)[i] += 1;
    SEXP cen = PROTECT(Rf_allocMatrix(REALSXP, (int)d.n, K));          /* d x K image of the K x d row-major centres */
    CHECK(scm_d2h(ctx(), REAL(cen), dc, (int64_t)K * d.n * 8));
    SEXP w = PROTECT(Rf_allocVector(REALSXP, K)), sz = PROTECT(Rf_allocVector(INTSXP, K));
    for (int c = 0; c < K; c++) { REAL(w)[c] = wss[c]; INTEGER(sz)[c] = (int)size[c]; }
    SET_VECTOR_ELT(out, 0, cl); SET_VECTOR_ELT(out, 1, cen); SET_VECTOR_ELT(out, 2, w); SET_VECTOR_ELT(out, 3, Rf_ScalarReal(tot));
    SET_VECTOR_ELT(out, 4, sz); SET_VECTOR_ELT(out, 5, Rf_ScalarInteger(info[1]));
    SET_VECTOR_ELT(out, 6, Rf_ScalarInteger((info[1] >= iter_max || info[3]) ? 2 : 0));
    scm_dev_free(ctx(), dl); scm_dev_free(ctx(), dc); scm_dev_free(ctx(), d.p);
    UNPROTECT(5);
    return out;
}

static const R_CallMethodDef calls[] = {
    {"scm_kmeans_", (DL_FUNC)&scm_kmeans_, 5}, {"scm_ruv1_", (DL_FUNC)&scm_ruv1_, 5}, {"scm_ruvg2", (DL_FUNC)&scm_ruvg2, 5},
    {"scm_centroid_sqdist", (DL_FUNC)&scm_centroid_sqdist, 3}, {"scm_dist", (DL_FUNC)&scm_dist, 5}, {"scm_pca", (DL_FUNC)&scm_pca, 2},
    {"scm_use_gpus", (DL_FUNC)&scm_use_gpus, 1}, {"scm_host", (DL_FUNC)&scm_host, 12},
    {"scm_upload_csc", (DL_FUNC)&scm_upload_csc, 5}, {"scm_csc_pseudobulk", (DL_FUNC)&scm_csc_pseudobulk, 3},
    {"scm_csc_adjust_", (DL_FUNC)&scm_csc_adjust_, 5},
    {"scm_upload", (DL_FUNC)&scm_upload, 2}, {"scm_fit", (DL_FUNC)&scm_fit, 8}, {"scm_fit_gram", (DL_FUNC)&scm_fit_gram, 9},
    {"scm_release_scratch", (DL_FUNC)&scm_release_scratch, 0},
    {"scm_adjust_", (DL_FUNC)&scm_adjust_, 8}, {"scm_segmean", (DL_FUNC)&scm_segmean, 3},
    {"scm_upload_sparse", (DL_FUNC)&scm_upload_sparse, 5}, {"scm_ruvg", (DL_FUNC)&scm_ruvg, 3}, {"scm_sil", (DL_FUNC)&scm_sil, 14},
    {NULL, NULL, 0}};
void R_init_scMergeB200(DllInfo* dll) { R_registerRoutines(dll, NULL, calls, NULL, NULL); R_useDynamicSymbols(dll, FALSE); }

// K6 (primary path): fused rank-k update with TMA-staged row tiles.
//
//   out = Y - ((Y - center) B) A          (same math and fragment tricks as adjust.cuh)
//
// The register-staged kernel in adjust.cuh can keep only ~1 block of loads per warp in flight, which leaves
// HBM half idle (long-scoreboard stalls, ncu profiles/r01).  Here the cell rows are streamed through a shared
// memory ring by the TMA unit instead:
//   * persistent CTAs (2 per SM), each walks row tiles of 8*MT cells; a producer warp issues
//     `cp.async.bulk.tensor.2d` boxes of [8*MT cells x 16 genes] (SWIZZLE_128B) into a ring of stages of 4 boxes
//     (one per consumer warp), running ahead across pass and tile boundaries,
//   * 4 consumer warps: warp w owns box w of every stage.  Pass A (W += Y_blk B_blk over the control-gene
//     blocks) and pass B (out_blk = Y_blk - W A_blk) read their DMMA fragments from the swizzled box with two
//     conflict-free LDS.128 per 8-cell m-tile; pass B writes the result straight to global memory
//     (8 full 128-byte lines per warp store),
//   * the small-operand fragments (one KT-KB block of B or A per box, L2 resident) are prefetched one box ahead
//     by the consumer warp itself with `cp.async` into a private double buffer: every lane copies exactly the
//     16-byte pieces it will read, so `cp.async.wait_group` is the only synchronisation needed.  (Staging them
//     with `cp.async.bulk` on the ring's mbarrier produced rare stale fragments -- ~2 % of launches had one wrong
//     16-gene block, tools/probe/stress_adjust.py -- and was dropped.)
//   * the four partial W tiles are combined through shared memory between the passes (named barrier of the 128
//     consumer threads, fixed order => deterministic),
//   * pass B re-reads from L2 what pass A fetched (one tile's control-gene columns per CTA: <= 8*MT*c*8 bytes),
//   * a ring slot is handed back to the producer only after the (volatile, ordered) DMMAs that consume its
//     shared-memory loads have issued.
// Needs ld % 2 == 0 and 16-byte aligned bases (TMA); otherwise, and for the subset / W-output variants, the
// register-staged kernel is used.
#pragma once
#include "adjust.cuh"
#include "gram_tma.cuh"

namespace scm {

constexpr int AT_CONSUMERS = 4;
constexpr int AT_THREADS = (AT_CONSUMERS + 1) * 32;  // 4 consumer warps + 1 producer warp

template <int KT, int MT>
struct AdjTile {
    static constexpr int ROWS = 8 * MT;
    static constexpr int BOX_BYTES = ROWS * 128;
    static constexpr int FRAG_BYTES = KT * 1024;  // one block of B (pass A) or A (pass B) fragments: 2*KT x 32 lanes x 16 B
    static constexpr int STAGE_BYTES = AT_CONSUMERS * BOX_BYTES;
    static constexpr int FRAGBUF_BYTES = AT_CONSUMERS * 2 * FRAG_BYTES;  // per-warp double buffer
    static constexpr int PART_DOUBLES = AT_CONSUMERS * MT * KT * 2 * 32;
    // ring depth chosen so that two CTAs (ring + fragment buffers + W partials) fit the 227 KB of an SM
    static constexpr int BUDGET = 112 * 1024 - 1024 - 256 - PART_DOUBLES * 8 - FRAGBUF_BYTES;
    static constexpr int STAGES = BUDGET / STAGE_BYTES >= 6 ? 6 : (BUDGET / STAGE_BYTES >= 2 ? BUDGET / STAGE_BYTES : 2);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 2 * STAGES * 8 + 64 + FRAGBUF_BYTES + PART_DOUBLES * 8 + 64;
};

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }

template <int KT, int MT>
__global__ void __launch_bounds__(AT_THREADS, 2)
adjust_tma_kernel(const __grid_constant__ CUtensorMap tmap, int64_t m, int64_t n, const double* __restrict__ Bfrag,
                  const double* __restrict__ Afrag, const double* __restrict__ w0, const int32_t* __restrict__ ctl_blocks,
                  int nblk_ctl, int nblk, int k, double* __restrict__ out, int64_t ldo, int vec_out) {
    using TL = AdjTile<KT, MT>;
    constexpr int AT_STAGES = TL::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by POINTER arithmetic: a round trip through uintptr_t hides the shared address space from the
    // compiler and every fragment load becomes a generic LD.E (64-bit addresses, long scoreboard) instead of LDS
    uint8_t* smem = smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + AT_STAGES * TL::STAGE_BYTES);
    uint64_t* empty = full + AT_STAGES;
    uint8_t* fragbuf = smem + AT_STAGES * TL::STAGE_BYTES + ((2 * AT_STAGES * 8 + 63) & ~63);
    double* s_part = reinterpret_cast<double*>(fragbuf + TL::FRAGBUF_BYTES);  // [4][MT*KT*2][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    if (tid == 0) {
        for (int s = 0; s < AT_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], AT_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    const int64_t ntiles = ceil_div(m, TL::ROWS);
    const int nA = (nblk_ctl + AT_CONSUMERS - 1) / AT_CONSUMERS, nB = (nblk + AT_CONSUMERS - 1) / AT_CONSUMERS;

    if (warp == AT_CONSUMERS) {
        // ===== producer: same (tile, pass, stage) walk as the consumers =====
        if (lane != 0) return;
        uint32_t g = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int row = (int)(t * TL::ROWS);
            for (int pass = 0; pass < 2; pass++) {
                const int ns = pass == 0 ? nA : nB, nb = pass == 0 ? nblk_ctl : nblk;
                for (int i = 0; i < ns; i++, g++) {
                    const int slot = g % AT_STAGES;
                    if (g >= AT_STAGES) mbar_wait(&empty[slot], ((g / AT_STAGES) - 1) & 1);
                    const int nbox = min(AT_CONSUMERS, nb - i * AT_CONSUMERS);
                    mbar_expect_tx(&full[slot], nbox * TL::BOX_BYTES);
                    for (int w = 0; w < nbox; w++) {
                        const int blk = pass == 0 ? ctl_blocks[i * AT_CONSUMERS + w] : i * AT_CONSUMERS + w;
                        tma_load_2d(smem + slot * TL::STAGE_BYTES + w * TL::BOX_BYTES, &tmap, blk * 16, row, &full[slot]);
                    }
                }
            }
        }
        return;
    }
    // ===== consumers =====
    // byte offsets of this lane's two 16-byte chunks (genes 4tig..4tig+1 and 4tig+2..4tig+3) inside an 8-row atom
    const int o0 = gid * 128 + (((2 * tig) ^ gid) << 4), o1 = gid * 128 + (((2 * tig + 1) ^ gid) << 4);
    // this warp's periodic fragment stream: cA pass-A blocks (B fragments) then cB pass-B blocks (A fragments) per tile
    const int cA = nblk_ctl > warp ? (nblk_ctl - warp + AT_CONSUMERS - 1) / AT_CONSUMERS : 0;
    const int cB = nblk > warp ? (nblk - warp + AT_CONSUMERS - 1) / AT_CONSUMERS : 0;
    uint8_t* myfrag = fragbuf + warp * 2 * TL::FRAG_BYTES + lane * 16;
    auto issue_frag = [&](int ord, int buf) {  // ord in [0, cA + cB): copy this lane's 2*KT pieces of that block
        const double* src = ord < cA ? Bfrag + (int64_t)ctl_blocks[warp + AT_CONSUMERS * ord] * KT * 128
                                     : Afrag + (int64_t)(warp + AT_CONSUMERS * (ord - cA)) * KT * 128;
#pragma unroll
        for (int q = 0; q < 2 * KT; q++) cp_async16(myfrag + buf * TL::FRAG_BYTES + q * 512, src + (q * 32 + lane) * 2, true);
        cp_async_commit();
    };
    int ford = 0, fbuf = 0;  // ordinal / buffer of the fragment block that is current (already issued)
    if (cA + cB > 0) issue_frag(0, 0);
    // returns this lane's fragment pointer for the current block and starts the copy of the next one
    auto take_frag = [&]() -> const double2* {
        cp_async_wait<0>();  // own pieces of the current block have landed (each lane reads only what it copied)
        const double2* p = reinterpret_cast<const double2*>(myfrag + fbuf * TL::FRAG_BYTES);
        ford = (ford + 1 == cA + cB) ? 0 : ford + 1;
        fbuf ^= 1;
        issue_frag(ford, fbuf);  // the buffer being overwritten was consumed (ordered DMMAs) one block ago
        return p;
    };

    uint32_t g = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t row0 = t * TL::ROWS;
        bool rowok[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) rowok[mt] = row0 + mt * 8 + gid < m;
        double w[MT][KT][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < KT; nt++) w[mt][nt][0] = w[mt][nt][1] = 0.0;
        // ---- pass A
        for (int i = 0; i < nA; i++, g++) {
            const int slot = g % AT_STAGES;
            mbar_wait(&full[slot], (g / AT_STAGES) & 1);
            if (i * AT_CONSUMERS + warp < nblk_ctl) {
                const uint8_t* box = smem + slot * TL::STAGE_BYTES + warp * TL::BOX_BYTES;
                double y[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const double2 a = *reinterpret_cast<const double2*>(box + mt * 1024 + o0);
                    const double2 b = *reinterpret_cast<const double2*>(box + mt * 1024 + o1);
                    y[mt][0] = a.x; y[mt][1] = a.y; y[mt][2] = b.x; y[mt][3] = b.y;
                }
                const double2* bf = take_frag();
#pragma unroll
                for (int nt = 0; nt < KT; nt++) {
                    const double2 b01 = bf[(nt * 2) * 32], b23 = bf[(nt * 2 + 1) * 32];
                    const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                    for (int j = 0; j < 4; j++)
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) dmma884_ordered(w[mt][nt][0], w[mt][nt][1], y[mt][j], b[j]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
        // ---- combine the partial W tiles
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < KT; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) s_part[(warp * (MT * KT * 2) + (mt * KT + nt) * 2 + e) * 32 + lane] = w[mt][nt][e];
        consumer_bar();
#pragma unroll
        for (int nt = 0; nt < KT; nt++) {
            const int kk = 8 * nt + 2 * tig;
            const double c0 = kk < k ? w0[kk] : 0.0, c1 = kk + 1 < k ? w0[kk + 1] : 0.0;
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                double t0 = 0.0, t1 = 0.0;
#pragma unroll
                for (int ww = 0; ww < AT_CONSUMERS; ww++) {
                    t0 += s_part[(ww * (MT * KT * 2) + (mt * KT + nt) * 2 + 0) * 32 + lane];
                    t1 += s_part[(ww * (MT * KT * 2) + (mt * KT + nt) * 2 + 1) * 32 + lane];
                }
                w[mt][nt][0] = -(t0 - c0); w[mt][nt][1] = -(t1 - c1);  // pass B adds (-W) A
            }
        }
        consumer_bar();  // s_part may be overwritten by the next tile from here on
        // ---- pass B
        double* outp[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) outp[mt] = out + (rowok[mt] ? row0 + mt * 8 + gid : 0) * ldo;
        for (int i = 0; i < nB; i++, g++) {
            const int slot = g % AT_STAGES;
            mbar_wait(&full[slot], (g / AT_STAGES) & 1);
            const int gb = i * AT_CONSUMERS + warp;
            if (gb < nblk) {
                const uint8_t* box = smem + slot * TL::STAGE_BYTES + warp * TL::BOX_BYTES;
                double y[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const double2 a = *reinterpret_cast<const double2*>(box + mt * 1024 + o0);
                    const double2 b = *reinterpret_cast<const double2*>(box + mt * 1024 + o1);
                    y[mt][0] = a.x; y[mt][1] = a.y; y[mt][2] = b.x; y[mt][3] = b.y;
                }
                const double2* af = take_frag();
                double2 a[2 * KT];
#pragma unroll
                for (int q = 0; q < 2 * KT; q++) a[q] = af[q * 32];
#pragma unroll
                for (int tau = 0; tau < 2; tau++)
#pragma unroll
                    for (int nt = 0; nt < KT; nt++)
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) {
                            dmma884_ordered(y[mt][2 * tau], y[mt][2 * tau + 1], w[mt][nt][0], a[tau * KT + nt].x);
                            dmma884_ordered(y[mt][2 * tau], y[mt][2 * tau + 1], w[mt][nt][1], a[tau * KT + nt].y);
                        }
                // hand the slot back only after the DMMAs that consume its loads have issued
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
                const int64_t g0 = (int64_t)gb * 16 + 4 * tig;
#pragma unroll
                for (int mt = 0; mt < MT; mt++) store4(outp[mt], g0, n, vec_out, rowok[mt], y[mt]);
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
            }
        }
    }
    cp_async_wait<0>();
}

template <int KT, int MT>
inline int adjust_tma_launch_t(scm_ctx* ctx, const CUtensorMap& tmap, int64_t m, int64_t n, const scm_coef* c, double* d_out,
                               int64_t ldo) {
    using TL = AdjTile<KT, MT>;
    const int smem = TL::SMEM_BYTES;
    SCM_CUDA(cudaFuncSetAttribute(adjust_tma_kernel<KT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t ntiles = ceil_div(m, TL::ROWS);
    int64_t grid = 2 * (int64_t)ctx->num_sms;
    if (grid > ntiles) grid = ntiles;
    const int vec_out = vec_flag(d_out, ldo);
    adjust_tma_kernel<KT, MT><<<(unsigned)grid, AT_THREADS, smem, ctx->stream>>>(tmap, m, n, c->Bfrag, c->Afrag, c->w0, c->ctl_blocks,
                                                                                 c->nblk_ctl, c->nblk, c->k, d_out, ldo, vec_out);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// Returns 1 when the TMA path is not applicable (caller uses the register-staged kernel), 0 on success, <0 on error.
inline int adjust_tma(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const scm_coef* c, double* d_out, int64_t ldo) {
    if (!ctx->use_tma || ld % 2 != 0 || ((uintptr_t)d_Y) % 16 != 0 || m >= ((int64_t)1 << 31) || n >= ((int64_t)1 << 31)) return 1;
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return 1;
    const int MT = c->KT <= 4 ? 4 : 2;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)m};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {16, (cuuint32_t)(8 * MT)};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)d_Y, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 1;
    switch (c->KT) {
        case 1: return adjust_tma_launch_t<1, 4>(ctx, tmap, m, n, c, d_out, ldo);
        case 2: return adjust_tma_launch_t<2, 4>(ctx, tmap, m, n, c, d_out, ldo);
        case 3: return adjust_tma_launch_t<3, 4>(ctx, tmap, m, n, c, d_out, ldo);
        case 4: return adjust_tma_launch_t<4, 4>(ctx, tmap, m, n, c, d_out, ldo);
        case 5: return adjust_tma_launch_t<5, 2>(ctx, tmap, m, n, c, d_out, ldo);
        case 6: return adjust_tma_launch_t<6, 2>(ctx, tmap, m, n, c, d_out, ldo);
        case 7: return adjust_tma_launch_t<7, 2>(ctx, tmap, m, n, c, d_out, ldo);
        case 8: return adjust_tma_launch_t<8, 2>(ctx, tmap, m, n, c, d_out, ldo);
        default: return 1;
    }
}

}  // namespace scm

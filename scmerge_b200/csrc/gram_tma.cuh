// K3 (primary path): Gram SYRK with TMA-staged operand tiles and an mbarrier full/empty pipeline.
//
// Same decomposition as gram_syrk.cuh (lower-triangular 128x128 tiles x S splits over cells, 8 warps x 64x32 warp
// tiles of DMMA.8x8x4, partial buffer + ordered reduce) but:
//   * operand tiles arrive by `cp.async.bulk.tensor.2d` (TMA): one elected thread issues 16 bulk copies per stage
//     ([16 cells x 16 genes] boxes, SWIZZLE_128B), so the compute warps execute no copy / address code at all;
//     out-of-range genes and cells are zero-filled by the TMA unit,
//   * stages are handed over with mbarriers (full: TMA transaction bytes, empty: one arrival per warp), there is
//     no __syncthreads in the main loop and warps may drift up to GT_STAGES-1 chunks apart, which keeps the
//     tensor pipe fed while another warp waits,
//   * fragment loads read the 128B-swizzled layout conflict-free by permuting the k index of each 8-cell
//     group (k-steps use cells {0,2,4,6} then {1,3,5,7}; any permutation of the summation index is legal),
//   * SHIFT = true: the column shift is subtracted in registers when fragments are loaded (12 DADD per 32 DMMA).  On sm_100
//     DADD and DMMA share the FP64 pipe and the DADDs sit in the DMMA issue stream: ncu's PC sampling puts 25 % of all samples
//     on them (math-pipe throttle) and the kernel runs 130.6 ms at 1M x 2000.  SHIFT = false (no shift vector: the caller hands
//     over cells that are already shifted, see ruv3_fit_impl) has no DADD at all and runs 117.3 ms = 34.1 TFLOP/s, 96 % of
//     cuBLAS DGEMM.  Shifting ONCE per stage in shared memory with the idle warps of the producer warpgroup was tried and is
//     slower (139.5 ms): next to two DMMA-saturating warps per scheduler a third warp's DADD waits ~110 cycles for the pipe
//     (80 % of the shifter warps' samples), so the shifters cannot keep up (profiles/README.md, r01i).
// Requires ld % 2 == 0 and a 16-byte aligned base (TMA global stride rule); otherwise gram_syrk.cuh is used.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gram_syrk.cuh"

namespace scm {

constexpr int GT_KC = 16;
constexpr int GT_STAGES = 6;
constexpr int GT_SLAB_BYTES = 16 * GT_KC * 8;      // one TMA box: 16 cells x 16 genes
constexpr int GT_TILE_BYTES = 8 * GT_SLAB_BYTES;   // 128 genes
constexpr int GT_STAGE_BYTES = 2 * GT_TILE_BYTES;  // A + B
constexpr int GT_SMEM_BYTES = GT_STAGES * GT_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}

// One 16-cell chunk: 4 k-steps x (8 + 4 fragment loads, 32 DMMA).  off[p][q] are the thread's swizzled byte offsets.
template <bool TAIL, bool SHIFT>
__device__ __forceinline__ void gram_chunk(const uint8_t* __restrict__ sa, const uint8_t* __restrict__ sb, const int (&off)[2][2],
                                           int wm, int wn, int tig, const double (&sA)[8], const double (&sB)[4], int rows_valid,
                                           double (&acc)[8][4][2]) {
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
        const int p = ks & 1;
        const int half = (ks >> 1) * 1024;  // second 8-cell swizzle atom of the box
        double af[8], bf[4];
#pragma unroll
        for (int a = 0; a < 8; a++)
            af[a] = *reinterpret_cast<const double*>(sa + (wm * 4 + (a >> 1)) * GT_SLAB_BYTES + half + off[p][a & 1]);
#pragma unroll
        for (int b = 0; b < 4; b++)
            bf[b] = *reinterpret_cast<const double*>(sb + (wn * 2 + (b >> 1)) * GT_SLAB_BYTES + half + off[p][b & 1]);
        if (!SHIFT) {
            // nothing to subtract; cells past the end of the matrix are zero-filled by the TMA unit
        } else if (TAIL) {
            const double msk = (8 * (ks >> 1) + 2 * tig + p) < rows_valid ? 1.0 : 0.0;  // zero-filled cells stay zero
#pragma unroll
            for (int a = 0; a < 8; a++) af[a] -= msk * sA[a];
#pragma unroll
            for (int b = 0; b < 4; b++) bf[b] -= msk * sB[b];
        } else {
#pragma unroll
            for (int a = 0; a < 8; a++) af[a] -= sA[a];
#pragma unroll
            for (int b = 0; b < 4; b++) bf[b] -= sB[b];
        }
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) dmma884_ordered(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
}

// Fragment of gene strip `t` (8 genes: rows t*8 .. t*8+7 of the tile) for k-step parity p / cell half `half`.
__device__ __forceinline__ double gram_frag(const uint8_t* __restrict__ tile, int t, int half, const int (&off)[2][2], int p) {
    return *reinterpret_cast<const double*>(tile + (t >> 1) * GT_SLAB_BYTES + half + off[p][t & 1]);
}

// THIN unit (last tile row when n is not a multiple of 128): only the first `na` 8-gene strips of the A side hold genes
// (na <= GT_THIN_NA), the rest is padding.  The 8 warps split the B side instead (2 strips each): na x 2 DMMA per k-step
// and warp instead of 8 x 4, i.e. the unit costs na/16 of a full one.
constexpr int GT_THIN_NA = 10;
template <bool TAIL, bool SHIFT>
__device__ __forceinline__ void gram_chunk_thin(const uint8_t* __restrict__ sa, const uint8_t* __restrict__ sb, const int (&off)[2][2],
                                                int warp, int tig, int na, const double (&sA)[GT_THIN_NA], const double (&sB)[2],
                                                int rows_valid, double (&acc)[GT_THIN_NA][2][2]) {
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
        const int p = ks & 1;
        const int half = (ks >> 1) * 1024;
        const double msk = !SHIFT ? 0.0 : (!TAIL || (8 * (ks >> 1) + 2 * tig + p) < rows_valid) ? 1.0 : 0.0;
        double af[GT_THIN_NA], bf[2];
#pragma unroll
        for (int a = 0; a < GT_THIN_NA; a++) af[a] = SHIFT ? gram_frag(sa, a, half, off, p) - msk * sA[a] : gram_frag(sa, a, half, off, p);  // strips >= na are zero padding
#pragma unroll
        for (int b = 0; b < 2; b++) bf[b] = SHIFT ? gram_frag(sb, 2 * warp + b, half, off, p) - msk * sB[b] : gram_frag(sb, 2 * warp + b, half, off, p);
#pragma unroll
        for (int a = 0; a < GT_THIN_NA; a++)
            if (a < na) {
#pragma unroll
                for (int b = 0; b < 2; b++) dmma884_ordered(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            }
    }
}

// DIAG unit (ti == tj): only the lower triangle is needed.  In 8-gene strips the triangle has 16 * 17 / 2 = 136 strip pairs
// instead of 256; warp w takes the row strips w and 15 - w ((w + 1) + (16 - w) = 17 pairs each, perfectly balanced), so a
// diagonal unit costs 17/32 of a full one.  `nvs` = number of strips that hold genes (16, less in the corner tile).
template <bool TAIL, bool SHIFT>
__device__ __forceinline__ void gram_chunk_diag(const uint8_t* __restrict__ sa, const int (&off)[2][2], int t0, int t1, int nvs, int tig,
                                                double s0, double s1, const double (&sS)[16], int rows_valid, double (&acc0)[16][2],
                                                double (&acc1)[16][2]) {
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
        const int p = ks & 1;
        const int half = (ks >> 1) * 1024;
        const double msk = !SHIFT ? 0.0 : (!TAIL || (8 * (ks >> 1) + 2 * tig + p) < rows_valid) ? 1.0 : 0.0;
        const double a0 = SHIFT ? gram_frag(sa, t0, half, off, p) - msk * s0 : gram_frag(sa, t0, half, off, p);
        const double a1 = SHIFT ? gram_frag(sa, t1, half, off, p) - msk * s1 : gram_frag(sa, t1, half, off, p);
        const bool v0 = t0 < nvs, v1 = t1 < nvs;
#pragma unroll
        for (int jb = 0; jb < 16; jb += 8) {  // two batches of 8 strips: all loads of a batch are issued before its DMMAs
            double f[8];
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] = SHIFT ? gram_frag(sa, jb + j, half, off, p) - msk * sS[jb + j] : gram_frag(sa, jb + j, half, off, p);  // always in bounds (padding strips are zero)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (v0 && jb + j <= t0) dmma884_ordered(acc0[jb + j][0], acc0[jb + j][1], a0, f[j]);
                if (v1 && jb + j <= t1) dmma884_ordered(acc1[jb + j][0], acc1[jb + j][1], a1, f[j]);
            }
        }
    }
}

// Unit order.  CTA c executes units c, c + grid, ... of a unit LIST.  With the split-major list (tiles of a split in triangular
// order) the unit classes (full / thin / diagonal) are interleaved: the lean units are cheaper than the full ones, CTAs drift
// up to ~1.5 units apart, the cells of a split are no longer shared through L2 and the DRAM traffic of the kernel is 165 GB at
// 1M x 2000 (10x the algorithmic 16 GB; 50 GB with full units only).  The list is therefore CLASS-ORDERED: [all full units]
// ++ [all thin] ++ [all diagonal], each split-major, so that the CTAs of a wave run units of equal cost on (nearly) the same
// cells and stay in step.  The list is a TABLE in global memory (gram_units_kernel, (split, tile) per entry): producer and
// consumers read their entry with one load where they used to divide by T, so the order costs no registers (round 1 computed
// the class-ordered mapping arithmetically in the consumer warps, which already fill their 232 registers, and spilled 60 bytes).
// 3 warpgroups: 8 consumer warps (setmaxnreg.inc -> 232 registers) + one producer warpgroup (setmaxnreg.dec -> 40;
// only its first lane works).  232 * 256 + 40 * 128 = 64512 <= 65536 registers.
constexpr int GT_THREADS = 384;

// Builds the class-ordered unit list (see "Unit order" above).  One CTA: tiles are classified once, then every thread fills its
// share of the S * T entries.  `lean` as in the kernel: bit 0 thin last tile row, bit 1 triangular diagonal units.
__global__ void __launch_bounds__(256)
gram_units_kernel(int ntile, int T, int S, int na_last, int lean, int2* __restrict__ list) {
    __shared__ int s_tile[3][1024];  // tiles of class 0 full / 1 thin / 2 diagonal, in triangular order
    __shared__ int s_cnt[3];
    if (threadIdx.x == 0) {
        int c[3] = {0, 0, 0};
        for (int t = 0; t < T && t < 1024; t++) {
            int ti, tj;
            tri_decode(t, ti, tj);
            int cls = 0;
            if ((lean & 2) && ti == tj) cls = 2;
            else if ((lean & 1) && ti != tj && ti == ntile - 1 && na_last <= GT_THIN_NA) cls = 1;
            s_tile[cls][c[cls]++] = t;
        }
        s_cnt[0] = c[0]; s_cnt[1] = c[1]; s_cnt[2] = c[2];
    }
    __syncthreads();
    const int n0 = s_cnt[0], n1 = s_cnt[1], n2 = s_cnt[2];
    const int64_t total = (int64_t)S * T;
    for (int64_t p = threadIdx.x; p < total; p += blockDim.x) {
        int cls, off;
        if (p < (int64_t)S * n0) { cls = 0; off = (int)p; }
        else if (p < (int64_t)S * (n0 + n1)) { cls = 1; off = (int)(p - (int64_t)S * n0); }
        else { cls = 2; off = (int)(p - (int64_t)S * (n0 + n1)); }
        const int nc = cls == 0 ? n0 : (cls == 1 ? n1 : n2);
        list[p] = make_int2(off / nc, s_tile[cls][off % nc]);
    }
}

template <bool SHIFT>
__global__ void __launch_bounds__(GT_THREADS, 1)
gram_tma_kernel(const __grid_constant__ CUtensorMap tmap, int64_t m, int64_t n, const double* __restrict__ shift, int T, int S,
                int64_t rows_per_split, double* __restrict__ partial, int ntile, int na_last, int lean,
                const int2* __restrict__ unit_list) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment.  By pointer arithmetic the compiler keeps the shared address space and emits LDS; through uintptr_t it
    // emits generic LD.E.  Measured on the SHIFT = true variant: generic 130.9 ms, LDS 132.6 ms (different instruction
    // schedule around the DADDs), so that variant keeps the generic form.
    uint8_t* smem = SHIFT ? reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023)
                          : smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + GT_STAGES * GT_STAGE_BYTES);
    uint64_t* empty = full + GT_STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    if (tid == 0) {
        for (int s = 0; s < GT_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    int off[2][2];
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
        for (int q = 0; q < 2; q++)
            off[p][q] = (2 * tig + p) * 128 + ((((gid >> 1) ^ (2 * tig + p)) ^ (q << 2)) << 4) + (gid & 1) * 8;

    const int64_t units = (int64_t)T * S;
    if (warp >= 8) {
        // ===== producer warpgroup: one lane walks the same unit / chunk sequence and keeps the ring full =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (tid != 256) return;
        uint32_t gl = 0;
        for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
            const int2 ud = unit_list ? unit_list[u] : make_int2((int)(u / T), (int)(u % T));  // (split, tile): see gram_units_kernel
            const int split = ud.x;
            int ti, tj;
            tri_decode(ud.y, ti, tj);
            const bool diag = (ti == tj);
            const int64_t row_begin = (int64_t)split * rows_per_split;
            const int64_t row_end = min(m, row_begin + rows_per_split);
            const int nchunk = row_end > row_begin ? (int)ceil_div(row_end - row_begin, GT_KC) : 0;
            const int colA0 = ti * GR_BT, colB0 = tj * GR_BT;
            for (int c = 0; c < nchunk; c++, gl++) {
                const int slot = gl % GT_STAGES;
                if (gl >= GT_STAGES) mbar_wait(&empty[slot], ((gl / GT_STAGES) - 1) & 1);
                uint8_t* sa = smem + slot * GT_STAGE_BYTES;
                mbar_expect_tx(&full[slot], diag ? GT_TILE_BYTES : GT_STAGE_BYTES);
                const int row = (int)(row_begin + (int64_t)c * GT_KC);
#pragma unroll
                for (int sl = 0; sl < 8; sl++) tma_load_2d(sa + sl * GT_SLAB_BYTES, &tmap, colA0 + 16 * sl, row, &full[slot]);
                if (!diag) {
#pragma unroll
                    for (int sl = 0; sl < 8; sl++)
                        tma_load_2d(sa + GT_TILE_BYTES + sl * GT_SLAB_BYTES, &tmap, colB0 + 16 * sl, row, &full[slot]);
                }
            }
        }
        return;
    }
    // ===== consumer warps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    uint32_t gu = 0;  // chunks consumed so far by this CTA
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const int2 ud = unit_list ? unit_list[u] : make_int2((int)(u / T), (int)(u % T));
        const int split = ud.x;
        int ti, tj;
        tri_decode(ud.y, ti, tj);
        const bool diag = (ti == tj);
        const int64_t row_begin = (int64_t)split * rows_per_split;
        const int64_t row_end = min(m, row_begin + rows_per_split);
        const int nchunk = row_end > row_begin ? (int)ceil_div(row_end - row_begin, GT_KC) : 0;
        const int colA0 = ti * GR_BT, colB0 = tj * GR_BT;
        double* out = partial + ((int64_t)split * T + ud.y) * (GR_BT * GR_BT);
        const bool last_row = ti == ntile - 1;
        auto shift_at = [&](int64_t g) -> double { return (SHIFT && shift && g < n) ? shift[g] : 0.0; };
        if ((lean & 2) && diag) {
            // ---- lower triangle only (see gram_chunk_diag)
            const int t0 = warp, t1 = 15 - warp, nvs = last_row ? na_last : 16;
            double sS[16], acc0[16][2], acc1[16][2];
#pragma unroll
            for (int j = 0; j < 16; j++) { sS[j] = shift_at(colA0 + j * 8 + gid); acc0[j][0] = acc0[j][1] = acc1[j][0] = acc1[j][1] = 0.0; }
            const double s0 = shift_at(colA0 + t0 * 8 + gid), s1 = shift_at(colA0 + t1 * 8 + gid);
            for (int c = 0; c < nchunk; c++) {
                const int slot = gu % GT_STAGES;
                mbar_wait(&full[slot], (gu / GT_STAGES) & 1);
                const uint8_t* sa = smem + slot * GT_STAGE_BYTES;
                const int64_t r0 = row_begin + (int64_t)c * GT_KC;
                if (r0 + GT_KC <= row_end) gram_chunk_diag<false, SHIFT>(sa, off, t0, t1, nvs, tig, s0, s1, sS, GT_KC, acc0, acc1);
                else gram_chunk_diag<SHIFT, SHIFT>(sa, off, t0, t1, nvs, tig, s0, s1, sS, (int)(row_end - r0), acc0, acc1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
                gu++;
            }
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (j <= t0) *reinterpret_cast<double2*>(out + (t0 * 8 + gid) * GR_BT + j * 8 + 2 * tig) = make_double2(acc0[j][0], acc0[j][1]);
                if (j <= t1) *reinterpret_cast<double2*>(out + (t1 * 8 + gid) * GR_BT + j * 8 + 2 * tig) = make_double2(acc1[j][0], acc1[j][1]);
            }
            continue;
        }
        if ((lean & 1) && !diag && last_row && na_last <= GT_THIN_NA) {
            // ---- padded last tile row: na_last strips x 2 strips per warp (see gram_chunk_thin)
            double sA[GT_THIN_NA], sB[2], acc[GT_THIN_NA][2][2];
#pragma unroll
            for (int a = 0; a < GT_THIN_NA; a++) { sA[a] = shift_at(colA0 + a * 8 + gid); acc[a][0][0] = acc[a][0][1] = acc[a][1][0] = acc[a][1][1] = 0.0; }
#pragma unroll
            for (int b = 0; b < 2; b++) sB[b] = shift_at(colB0 + (2 * warp + b) * 8 + gid);
            for (int c = 0; c < nchunk; c++) {
                const int slot = gu % GT_STAGES;
                mbar_wait(&full[slot], (gu / GT_STAGES) & 1);
                const uint8_t* sa = smem + slot * GT_STAGE_BYTES;
                const uint8_t* sb = sa + GT_TILE_BYTES;
                const int64_t r0 = row_begin + (int64_t)c * GT_KC;
                if (r0 + GT_KC <= row_end) gram_chunk_thin<false, SHIFT>(sa, sb, off, warp, tig, na_last, sA, sB, GT_KC, acc);
                else gram_chunk_thin<SHIFT, SHIFT>(sa, sb, off, warp, tig, na_last, sA, sB, (int)(row_end - r0), acc);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
                gu++;
            }
#pragma unroll
            for (int a = 0; a < GT_THIN_NA; a++)
                if (a < na_last) {
#pragma unroll
                    for (int b = 0; b < 2; b++)
                        *reinterpret_cast<double2*>(out + (a * 8 + gid) * GR_BT + (2 * warp + b) * 8 + 2 * tig) = make_double2(acc[a][b][0], acc[a][b][1]);
                }
            continue;
        }
        double sA[8], sB[4];
#pragma unroll
        for (int a = 0; a < 8; a++) sA[a] = shift_at(colA0 + wm * 64 + a * 8 + gid);
#pragma unroll
        for (int b = 0; b < 4; b++) sB[b] = shift_at(colB0 + wn * 32 + b * 8 + gid);
        double acc[8][4][2];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

        for (int c = 0; c < nchunk; c++) {
            const int slot = gu % GT_STAGES;
            mbar_wait(&full[slot], (gu / GT_STAGES) & 1);
            const uint8_t* sa = smem + slot * GT_STAGE_BYTES;
            const uint8_t* sb = diag ? sa : sa + GT_TILE_BYTES;
            const int64_t r0 = row_begin + (int64_t)c * GT_KC;
            if (r0 + GT_KC <= row_end) gram_chunk<false, SHIFT>(sa, sb, off, wm, wn, tig, sA, sB, GT_KC, acc);
            else gram_chunk<SHIFT, SHIFT>(sa, sb, off, wm, wn, tig, sA, sB, (int)(row_end - r0), acc);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            gu++;
        }
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int i = wm * 64 + a * 8 + gid, j = wn * 32 + b * 8 + 2 * tig;
                *reinterpret_cast<double2*>(out + i * GR_BT + j) = make_double2(acc[a][b][0], acc[a][b][1]);
            }
    }
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encoder() {
    static const PFN_tmapEncodeTiled fn = [] {  // C++11 static initialisation: safe when several device threads race here
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            return (PFN_tmapEncodeTiled)p;
        return (PFN_tmapEncodeTiled) nullptr;
    }();
    return fn;
}

// Returns 1 when the TMA path is not applicable (caller falls back to the cp.async kernel), 0 on success, <0 on error.
inline int gram_tma_launch(scm_ctx* ctx, Scratch& ws, const double* d_Y, int64_t m, int64_t n, int64_t ld, const double* d_shift,
                           const GramPlan& p, double* partial) {
    if (ld % 2 != 0 || ((uintptr_t)d_Y) % 16 != 0 || m >= ((int64_t)1 << 31) || n >= ((int64_t)1 << 31)) return 1;
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return 1;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)m};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {16, (cuuint32_t)GT_KC};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)d_Y, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1;
    SCM_CUDA(cudaFuncSetAttribute(gram_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM_BYTES));   // per device
    SCM_CUDA(cudaFuncSetAttribute(gram_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM_BYTES));
    const int valid_last = (int)(n - (int64_t)(p.ntile - 1) * GR_BT);
    const int na_last = (valid_last + 7) / 8;
    int2* units = nullptr;
    if (ctx->gram_class_order && ctx->gram_lean && p.T <= 1024 && p.S > 1) {
        SCM_TRY(ws.alloc(&units, (int64_t)p.S * p.T));
        gram_units_kernel<<<1, 256, 0, ctx->stream>>>(p.ntile, p.T, p.S, na_last, ctx->gram_lean, units);
        SCM_LAUNCH_CHECK(ctx);
    }
    if (d_shift)
        gram_tma_kernel<true><<<p.grid, GT_THREADS, GT_SMEM_BYTES, ctx->stream>>>(tmap, m, n, d_shift, p.T, p.S, p.rows_per_split, partial,
                                                                                  p.ntile, na_last, ctx->gram_lean, units);
    else
        gram_tma_kernel<false><<<p.grid, GT_THREADS, GT_SMEM_BYTES, ctx->stream>>>(tmap, m, n, nullptr, p.T, p.S, p.rows_per_split, partial,
                                                                                   p.ntile, na_last, ctx->gram_lean, units);
    return 0;
}

}  // namespace scm

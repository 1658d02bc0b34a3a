// liger_b200 -- the two sweeps over X as DENSE-TILE tensor-core contractions (tcgen05 + TMEM), for count-structured X.
//
// scaleData is  X[g, c] = N[g, c] * rs[g] * cs[c]  with integer counts N (normalize: cs = 1 / library size,
// R/preprocess.R:604; scaleNotCenter: rs = 1 / row RMS, src/data_processing.cpp:42-58).  Counts up to 256 are exact in
// BF16, and an fp32 factor entry is exactly the sum of three BF16 terms (8 + 8 + 8 mantissa bits), so
//
//   B[c, :] = cs[c] * sum_g N[g, c] * (rs[g] L~[g, :])          (rhs of the H solve, R/cINMF.R:480)
//   R[g, :] = rs[g] * sum_c N[g, c] * (cs[c] H[c, :])           (statistics of the V / W solves, R/cINMF.R:486-488,497-499)
//
// are products of an exact BF16 count tile with a 3-term BF16 split of the scaled factor: every product is exact and the
// accumulation is fp32 in TMEM -- the same arithmetic as the fp32 gather kernels (lgr_spmm.cu), whose shared-memory
// data path (one 128-byte row per non-zero) is what bounds them at ~20 % of the HBM roofline.  Here a non-zero costs one
// 2-byte shared-memory store (plus one to clear it), and the tensor core reads each 16-gene x 256-cell tile once.
//
// Both kernels are the same machine:   D[128 x 256] += A[128 x 16] * Xtile[256 x 16]^T   per k-step of 16,
//   rows of A / D  = (factor f, term t):  hi / mid / lo parts of the scaled factor (96 of the 128 rows are used),
//   k_dense_ltx:  columns = cells of a 256-cell tile, k = genes;  A slabs = pre-split rs * L~ (written by k_prep)
//   k_dense_xh:   columns = genes (2 accumulators = 512 genes per CTA), k = cells; A slabs = cs * H split in the kernel
// One persistent CTA per SM, warp-specialised:
//   scatter warps   stream the 4-byte entries (16-bit tile offset | bf16 count) of a stage and scatter them into a ring
//                   slot of shared memory in the canonical no-swizzle K-major UMMA layout; when the tensor core has
//                   consumed the slot (tcgen05.commit -> mbarrier) they clear exactly the entries they wrote
//   A producer      one TMA bulk copy per stage (k_dense_ltx) / converter warps splitting H rows (k_dense_xh)
//   MMA issuer      one thread: tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 256, K = 16, fp32 accumulators in TMEM
//   epilogue warps  tcgen05.ld, sum of the three terms, scaling, store of B rows / red.global.add into R
#include <cub/cub.cuh>

#include "lgr_common.cuh"
#include "lgr_ptx.cuh"

namespace lgr {

namespace {

constexpr int NST = 4;             // ring slots
constexpr int TILE_N = 256;        // columns of one accumulator (MMA N)
constexpr int KSTEP = 16;          // bf16 MMA K
constexpr int XTILE_BYTES = TILE_N * KSTEP * 2;   // 8 KB: one [256 x 16] bf16 operand tile
constexpr int ATILE_BYTES = 128 * KSTEP * 2;      // 4 KB: one [128 x 16] bf16 operand tile
constexpr int XSTAGE_BYTES = 4 * XTILE_BYTES;     // 32 KB per ring slot (both kernels)
constexpr int SCAT_WARPS = 3 * NST;               // three scatter warps per ring slot
constexpr int SCAT_GT = 96;                       // threads of one slot's scatter group
constexpr int PRE = 10;                           // entries per scatter thread held in registers across the slot wait

// ---- k_dense_ltx geometry ----
constexpr int K1_KS = 4;                          // k-steps per stage: 64 genes
constexpr int K1_ASTAGE_BYTES = K1_KS * ATILE_BYTES;   // 16 KB
constexpr int K1_WARPS = 6 + SCAT_WARPS;          // A producer, MMA, 4 epilogue, scatter
// ---- k_dense_xh geometry ----
constexpr int K2_KS = 2;                          // k-steps per stage: 32 cells
constexpr int K2_ASTAGE_BYTES = K2_KS * ATILE_BYTES;   // 8 KB
constexpr int K2_WARPS = 6 + SCAT_WARPS + NST;    // (spare), MMA, 4 epilogue, scatter, converters

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"l"(
                     (uint64_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start [0,14) >>4, LBO [16,30) >>4, SBO [32,46) >>4, version [46,48) = 1, no swizzle
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major, N = 256, M = 128
constexpr uint32_t IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(SCAT_GT) : "memory"); }

// x = hi + mid + lo exactly (each term a bf16; fp32 has 24 significant bits = 3 x 8)
__device__ __forceinline__ void split3(float x, unsigned short& hi, unsigned short& mid, unsigned short& lo) {
    const unsigned xb = __float_as_uint(x);
    const unsigned hb = xb & 0xFFFF0000u;
    const float r1 = x - __uint_as_float(hb);            // exact
    const unsigned mb = __float_as_uint(r1) & 0xFFFF0000u;
    const float r2 = r1 - __uint_as_float(mb);           // exact, at most 8 significant bits left
    hi = (unsigned short)(hb >> 16);
    mid = (unsigned short)(mb >> 16);
    lo = (unsigned short)(__float_as_uint(r2) >> 16);
}

// One slot's scatter group: clear the entries the slot held, then scatter the new stage's entries.
struct ScatterState {
    const uint32_t* old_ent = nullptr;
    int old_e0 = 0, old_e1 = 0;
};
__device__ __forceinline__ void scatter_stage(ScatterState& S, const uint32_t* __restrict__ ent, int e0, int e1, uint32_t xbase,
                                              uint64_t* empty_bar, uint32_t empty_parity, int gt, int bar_id) {
    uint32_t nv[PRE], ov[PRE];
#pragma unroll
    for (int j = 0; j < PRE; ++j) {   // both loads are issued before the wait: their latency overlaps the slot's MMAs
        const int en = e0 + gt + j * SCAT_GT, eo = S.old_e0 + gt + j * SCAT_GT;
        nv[j] = en < e1 ? __ldg(ent + en) : 0xFFFFFFFFu;
        ov[j] = eo < S.old_e1 ? __ldg(S.old_ent + eo) : 0xFFFFFFFFu;
    }
    mbar_wait(empty_bar, empty_parity);
#pragma unroll
    for (int j = 0; j < PRE; ++j)
        if (ov[j] != 0xFFFFFFFFu) sts_u16(xbase + ((ov[j] & 0xFFFFu) << 1), 0u);
    for (int e = S.old_e0 + gt + PRE * SCAT_GT; e < S.old_e1; e += SCAT_GT) sts_u16(xbase + ((__ldg(S.old_ent + e) & 0xFFFFu) << 1), 0u);
    group_bar(bar_id);   // every clear of this slot precedes every new store (an address may change owner)
#pragma unroll
    for (int j = 0; j < PRE; ++j)
        if (nv[j] != 0xFFFFFFFFu) sts_u16(xbase + ((nv[j] & 0xFFFFu) << 1), nv[j] >> 16);
    for (int e = e0 + gt + PRE * SCAT_GT; e < e1; e += SCAT_GT) {
        const uint32_t v = __ldg(ent + e);
        sts_u16(xbase + ((v & 0xFFFFu) << 1), v >> 16);
    }
    S.old_ent = ent;
    S.old_e0 = e0;
    S.old_e1 = e1;
}

struct SmemLayout {
    unsigned char* X;   // NST x 32 KB
    unsigned char* A;   // NST x a_stage
    uint64_t* fullX;    // NST
    uint64_t* fullA;    // NST
    uint64_t* empty;    // NST
    uint64_t* accfull;  // 2
    uint64_t* accempty; // 2
    uint32_t* tmem;
};
__device__ __forceinline__ SmemLayout carve(unsigned char* raw, int a_stage) {
    SmemLayout s;
    s.X = raw;
    s.A = raw + NST * XSTAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s.A + NST * a_stage);
    s.fullX = bars;
    s.fullA = bars + NST;
    s.empty = bars + 2 * NST;
    s.accfull = bars + 3 * NST;
    s.accempty = bars + 3 * NST + 2;
    s.tmem = reinterpret_cast<uint32_t*>(bars + 3 * NST + 4);
    return s;
}
size_t dense_smem(int a_stage) { return (size_t)NST * (XSTAGE_BYTES + a_stage) + (3 * NST + 4) * 8 + 16 + 1024; }

__device__ __forceinline__ unsigned char* align1k(unsigned char* p) {
    return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}

}  // namespace

// =====================================================================================================================
// k_dense_ltx:  B = L~^T X
// =====================================================================================================================
__global__ void __launch_bounds__(K1_WARPS * 32, 1) k_dense_ltx(const DenseDs* __restrict__ ds, int nds, int total_tiles, int dbg) {
    extern __shared__ unsigned char smem_dyn[];
    SmemLayout sm = carve(align1k(smem_dyn), K1_ASTAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup: barriers, TMEM (all 512 columns: two 256-column accumulators), cleared X ring ----
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(sm.fullX + s, SCAT_WARPS / NST);
            mbar_init(sm.fullA + s, 1);
            mbar_init(sm.empty + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(sm.accfull + a, 1);
            mbar_init(sm.accempty + a, 4);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        uint4* x4 = reinterpret_cast<uint4*>(sm.X);
        for (int e = threadIdx.x; e < NST * XSTAGE_BYTES / 16; e += blockDim.x) x4[e] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *sm.tmem;

    if (warp == 0) {
        // ===== A producer: one bulk copy (16 KB = four [128 x 16] slabs of rs * L~, pre-split by k_prep) per stage =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int i = find_ds(ds, nds, tile, &DenseDs::k1_tile_off);
                const DenseDs& D = ds[i];
                const unsigned char* src = reinterpret_cast<const unsigned char*>(D.Asplit);
                for (int s = 0; s < D.nstage1; ++s, ++it) {
                    const uint32_t slot = it % NST, ph = (it / NST) & 1u;
                    mbar_wait(sm.empty + slot, ph ^ 1u);
                    mbar_expect_tx(sm.fullA + slot, K1_ASTAGE_BYTES);
                    bulk_g2s(sm.A + slot * K1_ASTAGE_BYTES, src + (size_t)s * K1_ASTAGE_BYTES, K1_ASTAGE_BYTES, sm.fullA + slot);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
                const int i = find_ds(ds, nds, tile, &DenseDs::k1_tile_off);
                const int nstage = ds[i].nstage1;
                const uint32_t acc = tcount & 1u, accph = (tcount >> 1) & 1u;
                mbar_wait(sm.accempty + acc, accph ^ 1u);   // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem0 + acc * TILE_N;
                for (int s = 0; s < nstage; ++s, ++it) {
                    const uint32_t slot = it % NST, ph = (it / NST) & 1u;
                    mbar_wait(sm.fullA + slot, ph);
                    mbar_wait(sm.fullX + slot, ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sm.A + slot * K1_ASTAGE_BYTES), x0 = smem_u32(sm.X + slot * XSTAGE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < K1_KS; ++ks)
                        if (!(dbg & 2)) umma_bf16(tacc, umma_desc(a0 + ks * ATILE_BYTES, 128 * 16, 128), umma_desc(x0 + ks * XTILE_BYTES, TILE_N * 16, 128),
                                  IDESC_BF16, (s | ks) != 0);
                    umma_commit(sm.empty + slot);   // frees the slot (A slabs and X tiles) when these MMAs have read it
                }
                umma_commit(sm.accfull + acc);
            }
        }
        __syncwarp();
    } else if (warp < 6) {
        // ===== epilogue: D rows are 4 f + t (t = hi / mid / lo / zero): a warp's 32 TMEM lanes hold 8 factors x 4 terms =====
        const int q = warp & 3, t = lane & 3, fq = lane >> 2;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const int i = find_ds(ds, nds, tile, &DenseDs::k1_tile_off);
            const DenseDs& D = ds[i];
            const int c0 = (tile - D.k1_tile_off) * TILE_N;
            const int ncell = min(TILE_N, D.n - c0);
            const uint32_t acc = tcount & 1u, accph = (tcount >> 1) & 1u;
            mbar_wait(sm.accfull + acc, accph);
            tc_fence_after();
            float* __restrict__ Bout = D.B + (size_t)c0 * 32 + q * 8 + fq;
            const float* __restrict__ cs = D.cs + c0;
            for (int j = 0; j < TILE_N / 32; ++j) {
                if (j * 32 >= ncell) break;
                uint32_t v[32];
                tmem_ld32(tmem0 + ((uint32_t)(q * 32) << 16) + acc * TILE_N + j * 32, v);
#pragma unroll
                for (int c = 0; c < 32; ++c) {   // hi + mid + lo: the four lanes of a factor
                    float x = __uint_as_float(v[c]);
                    x += __shfl_xor_sync(0xffffffffu, x, 1);
                    x += __shfl_xor_sync(0xffffffffu, x, 2);
                    v[c] = __float_as_uint(x);
                }
#pragma unroll
                for (int g = 0; g < 8; ++g) {    // lane (fq, t) stores factor 8 q + fq of cell 4 g + t: 32-byte runs per cell
                    const uint32_t sel = t == 0 ? v[4 * g] : (t == 1 ? v[4 * g + 1] : (t == 2 ? v[4 * g + 2] : v[4 * g + 3]));
                    const int c = j * 32 + 4 * g + t;
                    if (c < ncell) Bout[(size_t)c * 32] = __uint_as_float(sel) * __ldg(cs + c);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.accempty + acc);
        }
    } else {
        // ===== scatter warps: slot = (warp - 6) % NST, three warps per slot =====
        const int sw = warp - 6, slot = sw % NST, gt = (sw / NST) * 32 + lane;
        const uint32_t xbase = smem_u32(sm.X + slot * XSTAGE_BYTES);
        ScatterState S;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int i = find_ds(ds, nds, tile, &DenseDs::k1_tile_off);
            const DenseDs& D = ds[i];
            const int* __restrict__ seg = D.seg1 + (size_t)(tile - D.k1_tile_off) * D.nstage1;
            for (int s = 0; s < D.nstage1; ++s, ++it) {
                if ((int)(it % NST) != slot) continue;
                const int e0 = __ldg(seg + s), e1 = (dbg & 1) ? e0 : __ldg(seg + s + 1);
                scatter_stage(S, D.ent1, e0, e1, xbase, sm.empty + slot, ((it / NST) & 1u) ^ 1u, gt, 1 + slot);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(sm.fullX + slot);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "n"(512) : "memory");
}

// =====================================================================================================================
// k_dense_xh:  R = X H     (flat work space: dataset x gene block of 512 x 32-cell stage, cut into equal ranges per CTA)
// =====================================================================================================================
struct K2Pos {
    int i;       // dataset
    int gb;      // gene block
    int cs0;     // first cell stage
    int cs1;     // end cell stage (exclusive) of this item inside the CTA's range
};
// the item that starts at flat position p (clipped to the CTA's range end pe)
__device__ __forceinline__ K2Pos k2_item(const DenseDs* __restrict__ ds, int nds, long long p, long long pe) {
    int i = 0;
    while (i + 1 < nds && p >= ds[i + 1].k2_flat_off) ++i;
    const long long local = p - ds[i].k2_flat_off;
    const int ncs = ds[i].ncstage;
    K2Pos r;
    r.i = i;
    r.gb = (int)(local / ncs);
    r.cs0 = (int)(local % ncs);
    const long long room = pe - p;
    r.cs1 = (int)min((long long)ncs, (long long)r.cs0 + room);
    return r;
}

__global__ void __launch_bounds__(K2_WARPS * 32, 1) k_dense_xh(const DenseDs* __restrict__ ds, int nds, long long total_flat, int dbg) {
    extern __shared__ unsigned char smem_dyn[];
    SmemLayout sm = carve(align1k(smem_dyn), K2_ASTAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long p0 = total_flat * blockIdx.x / gridDim.x, p1 = total_flat * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(sm.fullX + s, SCAT_WARPS / NST);
            mbar_init(sm.fullA + s, 1);
            mbar_init(sm.empty + s, 1);
        }
        mbar_init(sm.accfull, 1);
        mbar_init(sm.accempty, 4);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // X ring and A ring cleared once (rows 96..127 of every A slab stay zero for the whole kernel)
        uint4* x4 = reinterpret_cast<uint4*>(sm.X);
        for (int e = threadIdx.x; e < NST * (XSTAGE_BYTES + K2_ASTAGE_BYTES) / 16; e += blockDim.x) x4[e] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *sm.tmem;

    if (warp == 1) {
        // ===== MMA issuer: per k-step of 16 cells two MMAs (genes 0..255 and 256..511 of the block) sharing the A slab =====
        if (lane == 0) {
            uint32_t it = 0, icount = 0;
            for (long long p = p0; p < p1; ++icount) {
                const K2Pos P = k2_item(ds, nds, p, p1);
                mbar_wait(sm.accempty, (icount & 1u) ^ 1u);
                tc_fence_after();
                for (int cs = P.cs0; cs < P.cs1; ++cs, ++it) {
                    const uint32_t slot = it % NST, ph = (it / NST) & 1u;
                    mbar_wait(sm.fullA + slot, ph);
                    mbar_wait(sm.fullX + slot, ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sm.A + slot * K2_ASTAGE_BYTES), x0 = smem_u32(sm.X + slot * XSTAGE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < K2_KS; ++ks) {
                        const uint64_t ad = umma_desc(a0 + ks * ATILE_BYTES, 128 * 16, 128);
                        const uint32_t accum = (cs != P.cs0 || ks != 0) ? 1u : 0u;
                        if (dbg & 2) continue;
                        umma_bf16(tmem0, ad, umma_desc(x0 + (ks * 2 + 0) * XTILE_BYTES, TILE_N * 16, 128), IDESC_BF16, accum);
                        umma_bf16(tmem0 + TILE_N, ad, umma_desc(x0 + (ks * 2 + 1) * XTILE_BYTES, TILE_N * 16, 128), IDESC_BF16, accum);
                    }
                    umma_commit(sm.empty + slot);
                }
                umma_commit(sm.accfull);
                p += P.cs1 - P.cs0;
            }
        }
        __syncwarp();
    } else if (warp >= 2 && warp < 6) {
        // ===== epilogue: D rows are t * 32 + f, so quadrant q holds term q of all 32 factors; the three terms are summed by
        //       the reduction itself (red.global.add.f32 into R, which also sums over the CTAs that share a gene block) =====
        const int q = warp & 3;
        uint32_t icount = 0;
        for (long long p = p0; p < p1; ++icount) {
            const K2Pos P = k2_item(ds, nds, p, p1);
            const DenseDs& D = ds[P.i];
            mbar_wait(sm.accfull, icount & 1u);
            tc_fence_after();
            if (q < 3) {
                const int g0 = P.gb * 512;
                const int ngene = min(512, D.rows - g0);
                for (int j = 0; j < 512 / 32; ++j) {
                    if (j * 32 >= ngene) break;
                    uint32_t v[32];
                    tmem_ld32(tmem0 + ((uint32_t)(q * 32) << 16) + j * 32, v);
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int g = g0 + j * 32 + c;
                        if (j * 32 + c < ngene)
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(D.R + (size_t)g * 32 + lane),
                                         "f"(__uint_as_float(v[c]) * __ldg(D.rs + g))
                                         : "memory");
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.accempty);
            p += P.cs1 - P.cs0;
        }
    } else if (warp >= 6 && warp < 6 + SCAT_WARPS) {
        // ===== scatter warps =====
        const int sw = warp - 6, slot = sw % NST, gt = (sw / NST) * 32 + lane;
        const uint32_t xbase = smem_u32(sm.X + slot * XSTAGE_BYTES);
        ScatterState S;
        uint32_t it = 0;
        for (long long p = p0; p < p1;) {
            const K2Pos P = k2_item(ds, nds, p, p1);
            const DenseDs& D = ds[P.i];
            const int* __restrict__ seg = D.seg2 + (size_t)P.gb * D.ncstage;
            for (int cs = P.cs0; cs < P.cs1; ++cs, ++it) {
                if ((int)(it % NST) != slot) continue;
                const int e0 = __ldg(seg + cs), e1 = (dbg & 1) ? e0 : __ldg(seg + cs + 1);
                scatter_stage(S, D.ent2, e0, e1, xbase, sm.empty + slot, ((it / NST) & 1u) ^ 1u, gt, 1 + slot);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(sm.fullX + slot);
            }
            p += P.cs1 - P.cs0;
        }
    } else if (warp >= 6 + SCAT_WARPS) {
        // ===== converter warps (one per slot): A slab rows t * 32 + f = bf16 term t of cs[c] * H[c, f], 16 cells per k-step.
        //       Lane f reads H[c, f] for the stage's 32 cells (coalesced 128-byte rows) and writes one 16-byte chunk
        //       (8 cells) per term: consecutive lanes write consecutive chunks, no bank conflicts =====
        const int slot = warp - (6 + SCAT_WARPS);
        unsigned char* abase = sm.A + slot * K2_ASTAGE_BYTES;
        uint32_t it = 0;
        for (long long p = p0; p < p1;) {
            const K2Pos P = k2_item(ds, nds, p, p1);
            const DenseDs& D = ds[P.i];
            for (int cs = P.cs0; cs < P.cs1; ++cs, ++it) {
                if ((int)(it % NST) != slot) continue;
                const int c0 = cs * (K2_KS * KSTEP);
                float h[K2_KS * KSTEP];
#pragma unroll
                for (int j = 0; j < K2_KS * KSTEP; ++j) {
                    const int c = c0 + j;
                    h[j] = (c < D.n && !(dbg & 4)) ? __ldg(D.H + (size_t)c * 32 + lane) * __ldg(D.cs + c) : 0.f;
                }
                mbar_wait(sm.empty + slot, ((it / NST) & 1u) ^ 1u);
#pragma unroll
                for (int oc = 0; oc < K2_KS * 2; ++oc) {   // (k-step, chunk of 8 cells)
                    unsigned short hi[8], mid[8], lo[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) split3(h[oc * 8 + e], hi[e], mid[e], lo[e]);
                    uint4* dst = reinterpret_cast<uint4*>(abase + oc * 2048);
                    dst[lane] = make_uint4(hi[0] | (uint32_t)hi[1] << 16, hi[2] | (uint32_t)hi[3] << 16, hi[4] | (uint32_t)hi[5] << 16,
                                           hi[6] | (uint32_t)hi[7] << 16);
                    dst[32 + lane] = make_uint4(mid[0] | (uint32_t)mid[1] << 16, mid[2] | (uint32_t)mid[3] << 16,
                                                mid[4] | (uint32_t)mid[5] << 16, mid[6] | (uint32_t)mid[7] << 16);
                    dst[64 + lane] = make_uint4(lo[0] | (uint32_t)lo[1] << 16, lo[2] | (uint32_t)lo[3] << 16, lo[4] | (uint32_t)lo[5] << 16,
                                                lo[6] | (uint32_t)lo[7] << 16);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(sm.fullA + slot);
            }
            p += P.cs1 - P.cs0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "n"(512) : "memory");
}

// =====================================================================================================================
// G = H^T H  (k x k, fp64): one CTA per 1024-cell tile, 4 x 4 register patches of the upper triangle, fp64 atomics.
// (In the gather path this is phase 2 of k_spmm_xh; the dense R sweep does not stage fp32 H rows.)
// =====================================================================================================================
constexpr int GH_TILE = 1024;
constexpr int GH_THREADS = 288;     // 36 upper-triangular 4 x 4 patches (KP = 32) x 8 cell groups
__global__ void __launch_bounds__(GH_THREADS) k_gram_h(const DenseDs* __restrict__ ds, int nds, int k) {
    constexpr int KP = 32, NP = 8, NPATCH = 36, NCG = GH_THREADS / NPATCH;
    __shared__ float4 Hs[128 * NP];   // 128 cells x 32 floats at a time (16 KB)
    __shared__ float red[NCG * NPATCH * 16];
    const int i = find_ds(ds, nds, (int)blockIdx.x, &DenseDs::gh_blk_off);
    const DenseDs& D = ds[i];
    const int c0 = ((int)blockIdx.x - D.gh_blk_off) * GH_TILE;
    const int ncell = min(GH_TILE, D.n - c0);
    const int t = threadIdx.x, cg = t / NPATCH, pt = t % NPATCH;
    int pi = 0, pj = 0;
    {
        int rem = pt;
        while (rem >= NP - pi) {
            rem -= NP - pi;
            ++pi;
        }
        pj = pi + rem;
    }
    float a[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) a[x][y] = 0.f;
    const float4* __restrict__ H4 = reinterpret_cast<const float4*>(D.H + (size_t)c0 * KP);
    for (int b0 = 0; b0 < ncell; b0 += 128) {
        const int nb = min(128, ncell - b0);
        __syncthreads();
        for (int e = t; e < nb * NP; e += GH_THREADS) Hs[e] = __ldg(H4 + (size_t)b0 * NP + e);
        __syncthreads();
        for (int c = cg; c < nb; c += NCG) {
            const float4 u = Hs[c * NP + pi], v = Hs[c * NP + pj];
            const float uu[4] = {u.x, u.y, u.z, u.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) a[x][y] = fmaf(uu[x], vv[y], a[x][y]);
        }
    }
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) red[(cg * NPATCH + pt) * 16 + x * 4 + y] = a[x][y];
    __syncthreads();
    for (int e = t; e < NPATCH * 16; e += GH_THREADS) {
        const int p = e / 16, xy = e % 16, x = xy / 4, y = xy % 4;
        double sum = 0.0;
        for (int g = 0; g < NCG; ++g) sum += (double)red[(g * NPATCH + p) * 16 + xy];
        int ppi = 0, rem = p;
        while (rem >= NP - ppi) {
            rem -= NP - ppi;
            ++ppi;
        }
        const int ppj = ppi + rem;
        const int fa = ppi * 4 + x, fb = ppj * 4 + y;
        if (fa < k && fb < k) {
            atomicAdd(&D.G[fa * KP + fb], sum);
            if (ppi != ppj) atomicAdd(&D.G[fb * KP + fa], sum);
        }
    }
}

// =====================================================================================================================
// launches
// =====================================================================================================================
void launch_dense_ltx(const DenseDs* ds, int nds, int total_tiles, int num_sms, cudaStream_t st) {
    if (total_tiles <= 0) return;
    const int grid = std::min(total_tiles, num_sms);
    static const int dbg = getenv("LGR_DENSE_DBG") ? atoi(getenv("LGR_DENSE_DBG")) : 0;   // developer knob: 1 no scatter, 2 no MMA, 4 no H loads
    k_dense_ltx<<<grid, K1_WARPS * 32, dense_smem(K1_ASTAGE_BYTES), st>>>(ds, nds, total_tiles, dbg);
    LGR_CUDA(cudaGetLastError());
}
void launch_dense_xh(const DenseDs* ds, int nds, long long total_flat, int num_sms, cudaStream_t st) {
    if (total_flat <= 0) return;
    const int grid = (int)std::min<long long>(total_flat, num_sms);
    static const int dbg = getenv("LGR_DENSE_DBG") ? atoi(getenv("LGR_DENSE_DBG")) : 0;
    k_dense_xh<<<grid, K2_WARPS * 32, dense_smem(K2_ASTAGE_BYTES), st>>>(ds, nds, total_flat, dbg);
    LGR_CUDA(cudaGetLastError());
}
void launch_gram_h(const DenseDs* ds, int nds, int total_blocks, int k, cudaStream_t st) {
    if (total_blocks <= 0) return;
    k_gram_h<<<total_blocks, GH_THREADS, 0, st>>>(ds, nds, k);
    LGR_CUDA(cudaGetLastError());
}
int dense_gram_tile() { return GH_TILE; }
void dense_init() {
    LGR_CUDA(cudaFuncSetAttribute(k_dense_ltx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense_smem(K1_ASTAGE_BYTES)));
    LGR_CUDA(cudaFuncSetAttribute(k_dense_xh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense_smem(K2_ASTAGE_BYTES)));
}

// =====================================================================================================================
// one-time layout build (upload path): CSC + scales -> the two entry streams
//   ltx stream: segments (256-cell tile, 64-gene stage); entry offset addresses [k-step][chunk of 8 genes][cell][gene % 8]
//   xh  stream: segments (512-gene block, 32-cell stage); entry offset addresses [k-step][half block][chunk of 8 cells][gene][cell % 8]
// Every entry is verified to be an integer count in 1..256 (else *bad is raised: the scales do not factor X).
// =====================================================================================================================
__global__ void k_dense_keys(int n, int rows, const int* __restrict__ colptr, const int* __restrict__ rowidx, const float* __restrict__ val,
                             const float* __restrict__ rs, const float* __restrict__ cs, int nstage1, int ncstage,
                             unsigned int* __restrict__ key1, unsigned int* __restrict__ ent1, int* __restrict__ cnt1,
                             unsigned int* __restrict__ key2, unsigned int* __restrict__ ent2, int* __restrict__ cnt2, int* __restrict__ bad) {
    const int c = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (c >= n) return;
    const int b = colptr[c], e = colptr[c + 1];
    const float csc = cs[c];
    for (int t = b + lane; t < e; t += 32) {
        const int g = rowidx[t];
        const float q = val[t] / (rs[g] * csc);
        const float cntf = rintf(q);
        if (!(cntf >= 1.f && cntf <= 256.f && fabsf(q - cntf) <= 2e-3f)) atomicOr(bad, 1);
        const unsigned int bf = __float_as_uint(cntf) >> 16;   // exact: integers up to 256 have at most 8 significant bits
        {   // ltx: columns = cells of the tile, k = genes
            const int nn = c % TILE_N, kk = g % (K1_KS * KSTEP);
            const unsigned int off = (unsigned)((kk / KSTEP) * XTILE_BYTES + ((kk % KSTEP) / 8) * (TILE_N * 16) + nn * 16 + (kk % 8) * 2) >> 1;
            const unsigned int key = (unsigned)(c / TILE_N) * (unsigned)nstage1 + (unsigned)(g / (K1_KS * KSTEP));
            key1[t] = key;
            ent1[t] = off | (bf << 16);
            atomicAdd(&cnt1[key], 1);
        }
        {   // xh: columns = genes of the half block, k = cells
            const int gl = g % 512, nn = gl % TILE_N, sub = gl / TILE_N, kk = c % (K2_KS * KSTEP);
            const unsigned int off =
                (unsigned)(((kk / KSTEP) * 2 + sub) * XTILE_BYTES + ((kk % KSTEP) / 8) * (TILE_N * 16) + nn * 16 + (kk % 8) * 2) >> 1;
            const unsigned int key = (unsigned)(g / 512) * (unsigned)ncstage + (unsigned)(c / (K2_KS * KSTEP));
            key2[t] = key;
            ent2[t] = off | (bf << 16);
            atomicAdd(&cnt2[key], 1);
        }
    }
}

static void sort_stream(DBuf<unsigned int>& key, DBuf<unsigned int>& ent, DBuf<int>& cnt, size_t nnz, size_t nseg, DBuf<int>& seg,
                        DBuf<unsigned int>& out, cudaStream_t st) {
    DBuf<unsigned int> key2(nnz);
    out.alloc(std::max<size_t>(nnz, 1));
    seg.alloc(nseg + 1);
    int bits = 1;
    while (((size_t)1 << bits) < nseg) ++bits;
    size_t tmp_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, key2.p, ent.p, out.p, (int)nnz, 0, bits, st);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, cnt.p, seg.p, (int)(nseg + 1), st);
    DBuf<unsigned char> tmp(std::max(tmp_bytes, scan_bytes));
    LGR_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key.p, key2.p, ent.p, out.p, (int)nnz, 0, bits, st));
    LGR_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, cnt.p, seg.p, (int)(nseg + 1), st));
    LGR_CUDA(cudaStreamSynchronize(st));   // the temporaries are released on return
}

void dense_build(int n, int rows, const int* colptr, const int* rowidx, const float* val, size_t nnz, const float* rs, const float* cs,
                 DenseLayout& out, int* bad, cudaStream_t st) {
    out.nstage1 = (rows + K1_KS * KSTEP - 1) / (K1_KS * KSTEP);
    out.ntile1 = (n + TILE_N - 1) / TILE_N;
    out.ngb = (rows + 511) / 512;
    out.ncstage = (n + K2_KS * KSTEP - 1) / (K2_KS * KSTEP);
    const size_t nseg1 = (size_t)out.ntile1 * out.nstage1, nseg2 = (size_t)out.ngb * out.ncstage;
    LGR_REQUIRE(nseg1 < (size_t)0x7fffffff && nseg2 < (size_t)0x7fffffff && nnz < (size_t)0x7fffffff, "dense layout: too many segments");
    DBuf<unsigned int> key1(std::max<size_t>(nnz, 1)), ent1(std::max<size_t>(nnz, 1)), key2(std::max<size_t>(nnz, 1)), ent2(std::max<size_t>(nnz, 1));
    DBuf<int> cnt1(nseg1 + 1), cnt2(nseg2 + 1);
    cnt1.zero(st);
    cnt2.zero(st);
    if (nnz && n > 0) {
        k_dense_keys<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, st>>>(n, rows, colptr, rowidx, val, rs, cs, out.nstage1, out.ncstage,
                                                                                 key1.p, ent1.p, cnt1.p, key2.p, ent2.p, cnt2.p, bad);
        LGR_CUDA(cudaGetLastError());
    }
    sort_stream(key1, ent1, cnt1, nnz, nseg1, out.seg1, out.ent1, st);
    key1.release();
    ent1.release();
    sort_stream(key2, ent2, cnt2, nnz, nseg2, out.seg2, out.ent2, st);
    // pre-split factor image of k_dense_ltx: nstage1 x 16 KB, zero (rows of term 3 and rows >= `rows` stay zero)
    out.Asplit.alloc((size_t)out.nstage1 * K1_ASTAGE_BYTES / 2);
    out.Asplit.zero(st);
}

}  // namespace lgr

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
#ifndef LGR_SCAT_PER_SLOT
#define LGR_SCAT_PER_SLOT 3
#endif
constexpr int SCAT_PER_SLOT = LGR_SCAT_PER_SLOT;     // scatter warps per ring slot
constexpr int SCAT_WARPS = SCAT_PER_SLOT * NST;
constexpr int SCAT_GT = 32 * SCAT_PER_SLOT;       // threads of one slot's scatter group
constexpr int PRE = SCAT_PER_SLOT == 3 ? 10 : 8;  // entries per scatter thread held in registers (capacity PRE x SCAT_GT per stage)

// ---- k_dense_ltx geometry ----
constexpr int K1_KS = 4;                          // k-steps per stage: 64 genes
constexpr int K1_ASTAGE_BYTES = K1_KS * ATILE_BYTES;   // 16 KB
constexpr int K1_WARPS = 6 + SCAT_WARPS;          // A producer, MMA, 4 epilogue, scatter
// ---- k_dense_xh geometry ----
constexpr int K2_KS = 2;                          // k-steps per stage: 32 cells
constexpr int K2_ASTAGE_BYTES = K2_KS * ATILE_BYTES;   // 8 KB

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

constexpr uint32_t NOENT = 0xFFFFFFFFu;
__device__ unsigned long long g_dense_dbg[16];   // developer counters (LGR_DENSE_DBG & 16): cycles of CTA 0's MMA thread

// ---- stage iterators: the sequence of pipeline stages of this CTA (global index `it`) with their entry segments ----
// k_dense_ltx: tiles blockIdx.x, blockIdx.x + gridDim.x, ...; every tile walks the gene stages of its dataset
// The gene stages of a tile may be accumulated in any order: every CTA starts at its own rotation, so that the CTAs that
// run side by side do not all ask the L2 for the same 16 KB slab of the factor image at the same time.
__device__ __forceinline__ int k1_rot(int nstage) { return (int)((blockIdx.x * 5u) % (unsigned)nstage); }

struct K1Stages {
    const DenseDs* ds;
    int nds, total_tiles;
    int tile, s, nstage, rot;
    uint32_t it;
    const int* seg;
    const uint32_t* ent;
    __device__ __forceinline__ void init(const DenseDs* ds_, int nds_, int total_tiles_) {
        ds = ds_;
        nds = nds_;
        total_tiles = total_tiles_;
        tile = (int)blockIdx.x - (int)gridDim.x;
        s = 0;
        nstage = 0;
        rot = 0;
        it = 0xFFFFFFFFu;
        seg = nullptr;
        ent = nullptr;
    }
    __device__ __forceinline__ bool advance() {
        ++s;
        ++it;
        if (s >= nstage) {
            tile += (int)gridDim.x;
            if (tile >= total_tiles) return false;
            const int i = find_ds(ds, nds, tile, &DenseDs::k1_tile_off);
            const DenseDs& D = ds[i];
            s = 0;
            nstage = D.nstage1;
            rot = k1_rot(nstage);
            seg = D.seg1 + (size_t)(tile - D.k1_tile_off) * D.nstage1;
            ent = D.ent1;
        }
        return true;
    }
    __device__ __forceinline__ const int* segp() const {
        int sr = s + rot;
        if (sr >= nstage) sr -= nstage;
        return seg + sr;
    }
};
// k_dense_xh: flat positions [p0, p1) of (dataset, gene block, cell stage)
struct K2Stages {
    const DenseDs* ds;
    int nds;
    long long p, p1, item_end;
    int i, cs;
    uint32_t it;
    const int* seg;
    const uint32_t* ent;
    __device__ __forceinline__ void init(const DenseDs* ds_, int nds_, long long p0_, long long p1_) {
        ds = ds_;
        nds = nds_;
        p = p0_ - 1;
        p1 = p1_;
        item_end = p0_;
        i = 0;
        cs = 0;
        it = 0xFFFFFFFFu;
        seg = nullptr;
        ent = nullptr;
    }
    __device__ __forceinline__ bool advance() {
        ++p;
        ++it;
        ++cs;
        if (p >= p1) return false;
        if (p >= item_end) {
            i = 0;
            while (i + 1 < nds && p >= ds[i + 1].k2_flat_off) ++i;
            const DenseDs& D = ds[i];
            const long long local = p - D.k2_flat_off;
            const int gb = (int)(local / D.ncstage);
            cs = (int)(local % D.ncstage);
            item_end = D.k2_flat_off + (long long)(gb + 1) * D.ncstage;
            seg = D.seg2 + (size_t)gb * D.ncstage;
            ent = D.ent2;
        }
        return true;
    }
    __device__ __forceinline__ const int* segp() const { return seg + cs; }
};

// One slot's scatter group (96 threads), software-pipelined over its rounds (= the stages with it % NST == slot):
//   end of round r    the entries of round r + 2 are loaded into registers (their bounds were loaded at the end of round
//                     r - 1), the bounds of round r + 3 are loaded
//   round r           wait for the slot, clear the entries of the slot's previous stage (round r - 1, still in registers),
//                     group barrier, scatter the entries of round r, proxy fence, publish.
// A prefetch distance of TWO rounds: when the scatter side is the slower one a round lasts only as long as its own stores,
// which is less than a memory latency, and the proxy fence of a round waits for every load the thread has in flight.
// Three register arrays rotate through the roles (being scattered / to be cleared, then reloaded / in flight); the loop is
// unrolled three times so that no array is ever copied while its loads are in flight.
#ifndef LGR_PF_AHEAD
#define LGR_PF_AHEAD 13120
#endif
constexpr int PF_AHEAD = LGR_PF_AHEAD;   // entries (16 stages of ~820)
template <typename Stages>
__device__ __forceinline__ void scatter_role(Stages I, int slot, int gt, int lane, uint32_t xbase, uint64_t* empty_bar, uint64_t* full_bar,
                                             int dbg) {
    struct Round {
        const uint32_t* ent;
        const int* sp;
        int e0, e1;
        uint32_t it;
        bool valid;
    };
    auto locate = [&](Round& r) {            // next stage of this slot: index arithmetic on the shared-memory descriptors
        r.valid = false;
        while (I.advance()) {
            if ((int)(I.it % NST) == slot) {
                r.valid = true;
                break;
            }
        }
        r.ent = I.ent;
        r.it = I.it;
        r.sp = I.segp();
        r.e0 = r.e1 = 0;
    };
    auto load_bounds = [&](Round& r) {
        if (r.valid) {
            r.e0 = __ldg(r.sp);
            r.e1 = (dbg & 1) ? r.e0 : __ldg(r.sp + 1);
        }
    };
    auto load_entries = [&](const Round& r, uint32_t(&v)[PRE]) {
#pragma unroll
        for (int j = 0; j < PRE; ++j) {
            const int e = r.e0 + gt + j * SCAT_GT;
            v[j] = (r.valid && e < r.e1) ? __ldg(r.ent + e) : NOENT;
        }
    };
    Round cur, r1, r2, r3, old{nullptr, nullptr, 0, 0, 0, false};
    uint32_t E0[PRE], E1[PRE], E2[PRE];
#pragma unroll
    for (int j = 0; j < PRE; ++j) E2[j] = NOENT;
    locate(cur);
    load_bounds(cur);
    load_entries(cur, E0);
    locate(r1);
    load_bounds(r1);
    load_entries(r1, E1);
    locate(r2);
    load_bounds(r2);
    locate(r3);
    // one round: cv = entries of the current round, ov = entries of the previous round (cleared, then reloaded for round + 2)
    auto one_round = [&](uint32_t(&cv)[PRE], uint32_t(&ov)[PRE]) {
        mbar_wait(empty_bar, ((cur.it / NST) & 1u) ^ 1u);
#pragma unroll
        for (int j = 0; j < PRE; ++j)
            if (ov[j] != NOENT) sts_u16(xbase + ((ov[j] & 0xFFFFu) << 1), 0u);
        if (old.valid)
            for (int e = old.e0 + gt + PRE * SCAT_GT; e < old.e1; e += SCAT_GT) sts_u16(xbase + ((__ldg(old.ent + e) & 0xFFFFu) << 1), 0u);
        group_bar(1 + slot);   // every clear of this slot precedes every new store (an address may change owner)
#pragma unroll
        for (int j = 0; j < PRE; ++j)
            if (cv[j] != NOENT) sts_u16(xbase + ((cv[j] & 0xFFFFu) << 1), cv[j] >> 16);
        for (int e = cur.e0 + gt + PRE * SCAT_GT; e < cur.e1; e += SCAT_GT) {
            const uint32_t v = __ldg(cur.ent + e);
            sts_u16(xbase + ((v & 0xFFFFu) << 1), v >> 16);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar);
        // L2 prefetch of the entry stream PF_AHEAD entries further on (the stages of a tile / of a CTA's range are contiguous
        // in memory): the proxy fence of the NEXT round waits for the demand loads issued just below, so what a round costs is
        // their latency -- an L2 hit instead of a DRAM access.  A hint only: the entry arrays carry PF_AHEAD + 4 KB of slack.
        if ((dbg >> 8) > 0 && gt < 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(cur.ent + cur.e0 + ((dbg >> 8) << 8) + gt * 32));
        load_entries(r2, ov);   // round + 2 into the array that was just cleared from
        load_bounds(r3);        // round + 3
        old = cur;
        cur = r1;
        r1 = r2;
        r2 = r3;
        locate(r3);
    };
    for (;;) {
        if (!cur.valid) break;
        one_round(E0, E2);
        if (!cur.valid) break;
        one_round(E1, E0);
        if (!cur.valid) break;
        one_round(E2, E1);
    }
}

__device__ __forceinline__ unsigned char* align1k(unsigned char* p) {
    return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}

constexpr int MAX_SDS = 32;   // datasets whose descriptors are cached in shared memory (more: read from global memory)
constexpr int K1_NSA = 5;   // A-slab ring of k_dense_ltx (decoupled from the X ring: the slabs are fetched further ahead)
constexpr size_t K1_SMEM = (size_t)NST * XSTAGE_BYTES + (size_t)K1_NSA * K1_ASTAGE_BYTES + 40 * 8 + 1024;
constexpr int K2_NSA = 8;   // A-slab ring of k_dense_xh: the converters run up to 8 stages ahead of the MMAs
constexpr int K2_EPI_WORDS = 2 * 3 * 32 * 33;   // epilogue staging: two buffers x three terms x [32 factors][33]
constexpr size_t K2_SMEM = (size_t)NST * XSTAGE_BYTES + (size_t)K2_NSA * K2_ASTAGE_BYTES + (size_t)K2_EPI_WORDS * 4 + 40 * 8 + 1024;

// one lane of a converged warp (the MMA warps run their loops warp-uniformly, so that descriptors and addresses stay in
// uniform registers, and elect a lane only for the tcgen05 instructions themselves)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// non-blocking probe of an mbarrier phase (the MMA thread peeks at the next stage while the tensor core works)
__device__ __forceinline__ uint32_t mbar_test(uint64_t* bar, uint32_t phase) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return done;
}

// the dataset descriptors, cached in shared memory by the whole CTA (every role looks them up per tile / item)
__device__ __forceinline__ const DenseDs* cache_ds(const DenseDs* __restrict__ ds, int nds, DenseDs* sds) {
    if (nds > MAX_SDS) return ds;
    const int words = nds * (int)(sizeof(DenseDs) / 4);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(ds);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sds);
    for (int e = threadIdx.x; e < words; e += blockDim.x) dst[e] = __ldg(src + e);
    return sds;
}

}  // namespace

// =====================================================================================================================
// k_dense_ltx:  B = L~^T X
// X ring: NST slots of 32 KB (four [256 x 16] count tiles, written by the slot's three scatter warps; `full` = their three
// arrivals, `empty` = tcgen05.commit).  A ring: K1_NSA slots of 16 KB (four pre-split factor slabs, one bulk copy each;
// `fullA` = the copy's transaction bytes, `emptyA` = tcgen05.commit); it runs further ahead than the X ring.  The slabs are
// contiguous images written by k_prep, so the copies are 1-D bulk copies (cp.async.bulk, UBLKCP), not tensor-map loads.
// =====================================================================================================================
// LDF: row stride of B (= KP: 32, or 64 when the sweep runs once per half of the factors) as a compile-time constant -- the
// epilogue's address arithmetic shares its issue slots with the scatter warps
template <int LDF>
__global__ void __launch_bounds__(K1_WARPS * 32, 1) k_dense_ltx(const DenseDs* __restrict__ ds_g, int nds, int total_tiles, int dbg) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ DenseDs sds[MAX_SDS];
    unsigned char* raw = align1k(smem_dyn);
    unsigned char* smX = raw;
    unsigned char* smA = raw + NST * XSTAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smA + K1_NSA * K1_ASTAGE_BYTES);
    uint64_t* full = bars;              // NST: the three scatter warps of the X slot
    uint64_t* empty = bars + NST;       // NST
    uint64_t* fullA = bars + 2 * NST;   // K1_NSA: the producer's transaction bytes
    uint64_t* emptyA = fullA + K1_NSA;  // K1_NSA
    uint64_t* accfull = emptyA + K1_NSA; // 2
    uint64_t* accempty = accfull + 2;   // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup: barriers, TMEM (all 512 columns: two 256-column accumulators), cleared X ring ----
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(full + s, SCAT_WARPS / NST);
            mbar_init(empty + s, 1);
        }
        for (int s = 0; s < K1_NSA; ++s) {
            mbar_init(fullA + s, 1);
            mbar_init(emptyA + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(accfull + a, 1);
            mbar_init(accempty + a, 4);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        uint4* x4 = reinterpret_cast<uint4*>(smX);
        for (int e = threadIdx.x; e < NST * XSTAGE_BYTES / 16; e += blockDim.x) x4[e] = make_uint4(0u, 0u, 0u, 0u);
    }
    const DenseDs* ds = cache_ds(ds_g, nds, sds);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *tmem_slot;

    if (warp == 0) {
        // ===== A producer: one bulk copy (16 KB = four [128 x 16] slabs of rs * L~, pre-split by k_prep) per stage =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int i = find_ds(ds, nds, tile, &DenseDs::k1_tile_off);
                const DenseDs& D = ds[i];
                const unsigned char* src = reinterpret_cast<const unsigned char*>(D.Asplit);
                const int nstage = D.nstage1, rot = k1_rot(nstage);
                for (int s = 0; s < nstage; ++s, ++it) {
                    const uint32_t slot = it % K1_NSA, ph = (it / K1_NSA) & 1u;
                    int sr = s + rot;
                    if (sr >= nstage) sr -= nstage;
                    mbar_wait(emptyA + slot, ph ^ 1u);
                    if (dbg & 8) {
                        mbar_arrive(fullA + slot);
                        continue;
                    }
                    mbar_expect_tx(fullA + slot, K1_ASTAGE_BYTES);
                    bulk_g2s(smA + slot * K1_ASTAGE_BYTES, src + (size_t)sr * K1_ASTAGE_BYTES, K1_ASTAGE_BYTES, fullA + slot);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer.  The whole warp runs the loop (warp-uniform control flow keeps descriptors and barrier addresses in
        //       uniform registers); one elected lane issues the tcgen05 instructions.  One wait (peeked one stage ahead) and
        //       one commit per stage =====
        {
            const bool leader = elect_one();
            uint32_t it = 0, tcount = 0;
            long long wAcc = 0, wFull = 0, wIssue = 0, tstart = 0, c0 = 0, c1 = 0;
            const bool prof = (dbg & 16) != 0;
            if (prof) tstart = clock64();
            uint32_t ready = 0, readyA = 0, slot = 0, ph = 0, sa = 0, pha = 0;
            // descriptors of slot 0 / k-step 0; the address field (bits 0..13, in 16-byte units) is advanced by plain adds
            const uint64_t adesc0 = umma_desc(smem_u32(smA), 128 * 16, 128), xdesc0 = umma_desc(smem_u32(smX), TILE_N * 16, 128);
            // The tensor core queues only ~2 instructions: the issue of the 3rd MMA of a stage blocks until the 1st has
            // finished.  The probe of the next stage's barrier and the address arithmetic therefore sit between the 2nd and the
            // 3rd MMA, where they are hidden behind the queued work.
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
                const int i = find_ds(ds, nds, tile, &DenseDs::k1_tile_off);
                const int nstage = ds[i].nstage1;
                const uint32_t acc = tcount & 1u, accph = (tcount >> 1) & 1u;
                if (prof) c0 = clock64();
                mbar_wait(accempty + acc, accph ^ 1u);   // the epilogue has drained this accumulator
                if (prof) wAcc += clock64() - c0;
                const uint32_t tacc = tmem0 + acc * TILE_N;
                for (int s = 0; s < nstage; ++s, ++it) {
                    if (prof) c0 = clock64();
                    if (!readyA) mbar_wait(fullA + sa, pha);
                    if (!ready) mbar_wait(full + slot, ph);
                    if (prof) c1 = clock64(), wFull += c1 - c0;
                    tc_fence_after();
                    const uint64_t ad = adesc0 + (uint64_t)(sa * (K1_ASTAGE_BYTES >> 4)), xd = xdesc0 + (uint64_t)(slot * (XSTAGE_BYTES >> 4));
                    const bool mma = !(dbg & 2);
                    if (mma && leader) {
                        umma_bf16(tacc, ad, xd, IDESC_BF16, s != 0);
                        umma_bf16(tacc, ad + (uint64_t)(ATILE_BYTES >> 4), xd + (uint64_t)(XTILE_BYTES >> 4), IDESC_BF16, 1u);
                    }
                    const uint32_t ns = slot + 1 == NST ? 0u : slot + 1, nph = ns == 0 ? ph ^ 1u : ph;
                    const uint32_t nsa = sa + 1 == K1_NSA ? 0u : sa + 1, npha = nsa == 0 ? pha ^ 1u : pha;
                    ready = mbar_test(full + ns, nph);   // peek: consumed at the next stage (test_wait has ~150 cycles of latency)
                    readyA = mbar_test(fullA + nsa, npha);
                    if (mma && leader) {
                        umma_bf16(tacc, ad + (uint64_t)(2 * (ATILE_BYTES >> 4)), xd + (uint64_t)(2 * (XTILE_BYTES >> 4)), IDESC_BF16, 1u);
                        umma_bf16(tacc, ad + (uint64_t)(3 * (ATILE_BYTES >> 4)), xd + (uint64_t)(3 * (XTILE_BYTES >> 4)), IDESC_BF16, 1u);
                    }
                    if (leader) {   // frees the X slot and the A slot when these MMAs have read them
                        umma_commit(empty + slot);
                        umma_commit(emptyA + sa);
                    }
                    if (prof) wIssue += clock64() - c1;
                    slot = ns;
                    ph = nph;
                    sa = nsa;
                    pha = npha;
                }
                if (leader) umma_commit(accfull + acc);   // the tile's accumulator is complete: the epilogue may start
                __syncwarp();
            }
            if (prof && blockIdx.x == 0 && leader) {
                g_dense_dbg[0] = (unsigned long long)(clock64() - tstart);
                g_dense_dbg[1] = (unsigned long long)wAcc;
                g_dense_dbg[2] = (unsigned long long)wFull;
                g_dense_dbg[3] = 0;
                g_dense_dbg[4] = (unsigned long long)wIssue;
                g_dense_dbg[5] = 0;
                g_dense_dbg[6] = it;
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
            constexpr int ldf = LDF;
            float* __restrict__ Bout = D.B + (size_t)c0 * ldf + q * 8 + fq;
            const float* __restrict__ cs = D.cs + c0;
            mbar_wait(accfull + acc, accph);
            tc_fence_after();
            for (int j = 0; j < TILE_N / 32; ++j) {
                if (j * 32 >= ncell) break;
                float csv[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int c = j * 32 + 4 * g + t;
                    csv[g] = c < ncell ? __ldg(cs + c) : 0.f;
                }
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
                    if (c < ncell) Bout[(size_t)c * ldf] = __uint_as_float(sel) * csv[g];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accempty + acc);
        }
    } else {
        // ===== scatter warps: slot = (warp - 6) % NST, three warps per slot =====
        const int sw = warp - 6, slot = sw % NST, gt = (sw / NST) * 32 + lane;
        K1Stages I;
        I.init(ds, nds, total_tiles);
        scatter_role(I, slot, gt, lane, smem_u32(smX + slot * XSTAGE_BYTES), empty + slot, full + slot, dbg);
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

constexpr int K2_WARPS_TOTAL = 6 + SCAT_WARPS;   // A producer, MMA, 4 epilogue, scatter
constexpr int K2_IMG_STAGE_BYTES = K2_KS * 2 * 96 * 16;   // 6 KB: the 96 used rows of the four [128 x 8] chunks of a stage

template <int LDF>
__global__ void __launch_bounds__(K2_WARPS_TOTAL * 32, 1) k_dense_xh(const DenseDs* __restrict__ ds_g, int nds, long long total_flat, int dbg) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ DenseDs sds[MAX_SDS];
    unsigned char* raw = align1k(smem_dyn);
    unsigned char* smX = raw;
    unsigned char* smA = raw + NST * XSTAGE_BYTES;
    float* smEpi = reinterpret_cast<float*>(smA + K2_NSA * K2_ASTAGE_BYTES);   // epilogue staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(smEpi + K2_EPI_WORDS);
    uint64_t* full = bars;              // NST: the three scatter warps of the X slot
    uint64_t* empty = bars + NST;       // NST
    uint64_t* fullA = bars + 2 * NST;   // K2_NSA: the converter warp of the A slot
    uint64_t* emptyA = fullA + K2_NSA;  // K2_NSA
    uint64_t* accfull = emptyA + K2_NSA;
    uint64_t* accempty = accfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long p0 = total_flat * blockIdx.x / gridDim.x, p1 = total_flat * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(full + s, SCAT_WARPS / NST);
            mbar_init(empty + s, 1);
        }
        for (int s = 0; s < K2_NSA; ++s) {
            mbar_init(fullA + s, 1);   // the producer's arrive.expect_tx
            mbar_init(emptyA + s, 1);
        }
        mbar_init(accfull, 1);
        mbar_init(accempty, 4);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // X ring and A ring cleared once (rows 96..127 of every A slab stay zero for the whole kernel)
        uint4* x4 = reinterpret_cast<uint4*>(smX);
        for (int e = threadIdx.x; e < (NST * XSTAGE_BYTES + K2_NSA * K2_ASTAGE_BYTES) / 16; e += blockDim.x) x4[e] = make_uint4(0u, 0u, 0u, 0u);
    }
    const DenseDs* ds = cache_ds(ds_g, nds, sds);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *tmem_slot;

    if (warp == 0) {
        // ===== A producer: the cs-scaled bf16 split of H is written by k_gram_tc as an image of the operand slabs (without
        //       the 32 padding rows of each [128 x 8] chunk); four bulk copies of 1536 B per stage, 8 stages ahead =====
        if (lane == 0) {
            uint32_t it = 0;
            for (long long p = p0; p < p1;) {
                const K2Pos P = k2_item(ds, nds, p, p1);
                const unsigned char* src = reinterpret_cast<const unsigned char*>(ds[P.i].Hs);
                for (int cs = P.cs0; cs < P.cs1; ++cs, ++it) {
                    const uint32_t sa = it % K2_NSA, pha = (it / K2_NSA) & 1u;
                    mbar_wait(emptyA + sa, pha ^ 1u);
                    mbar_expect_tx(fullA + sa, K2_IMG_STAGE_BYTES);
#pragma unroll
                    for (int oc = 0; oc < K2_KS * 2; ++oc)
                        bulk_g2s(smA + sa * K2_ASTAGE_BYTES + oc * 2048, src + (size_t)cs * K2_IMG_STAGE_BYTES + oc * 1536, 1536, fullA + sa);
                }
                p += P.cs1 - P.cs0;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (converged warp, elected lane): per k-step of 16 cells two MMAs (genes 0..255 and 256..511 of the block)
        //       sharing the A slab =====
        {
            const bool leader = elect_one();
            uint32_t it = 0, icount = 0;
            long long wAcc = 0, wFull = 0, wIssue = 0, tstart = 0, c0 = 0, c1 = 0;
            const bool prof = (dbg & 16) != 0;
            if (prof) tstart = clock64();
            uint32_t ready = 0, readyA = 0, slot = 0, ph = 0, sa = 0, pha = 0;
            const uint64_t adesc0 = umma_desc(smem_u32(smA), 128 * 16, 128), xdesc0 = umma_desc(smem_u32(smX), TILE_N * 16, 128);
            for (long long p = p0; p < p1; ++icount) {
                const K2Pos P = k2_item(ds, nds, p, p1);
                if (prof) c0 = clock64();
                mbar_wait(accempty, (icount & 1u) ^ 1u);
                if (prof) wAcc += clock64() - c0;
                for (int cs = P.cs0; cs < P.cs1; ++cs, ++it) {
                    if (prof) c0 = clock64();
                    if (!readyA) mbar_wait(fullA + sa, pha);
                    if (!ready) mbar_wait(full + slot, ph);
                    if (prof) c1 = clock64(), wFull += c1 - c0;
                    tc_fence_after();
                    const uint64_t ad = adesc0 + (uint64_t)(sa * (K2_ASTAGE_BYTES >> 4)), xd = xdesc0 + (uint64_t)(slot * (XSTAGE_BYTES >> 4));
                    const uint32_t accum = cs != P.cs0 ? 1u : 0u;
                    const bool mma = !(dbg & 2);
                    if (mma && leader) {   // k-step 0
                        umma_bf16(tmem0, ad, xd, IDESC_BF16, accum);
                        umma_bf16(tmem0 + TILE_N, ad, xd + (uint64_t)(XTILE_BYTES >> 4), IDESC_BF16, accum);
                    }
                    const uint32_t ns = slot + 1 == NST ? 0u : slot + 1, nph = ns == 0 ? ph ^ 1u : ph;
                    const uint32_t nsa = sa + 1 == K2_NSA ? 0u : sa + 1, npha = nsa == 0 ? pha ^ 1u : pha;
                    ready = mbar_test(full + ns, nph);
                    readyA = mbar_test(fullA + nsa, npha);
                    if (mma && leader) {   // k-step 1
                        umma_bf16(tmem0, ad + (uint64_t)(ATILE_BYTES >> 4), xd + (uint64_t)(2 * (XTILE_BYTES >> 4)), IDESC_BF16, 1u);
                        umma_bf16(tmem0 + TILE_N, ad + (uint64_t)(ATILE_BYTES >> 4), xd + (uint64_t)(3 * (XTILE_BYTES >> 4)), IDESC_BF16, 1u);
                    }
                    if (leader) {
                        umma_commit(empty + slot);
                        umma_commit(emptyA + sa);
                    }
                    if (prof) wIssue += clock64() - c1;
                    slot = ns;
                    ph = nph;
                    sa = nsa;
                    pha = npha;
                }
                if (leader) umma_commit(accfull);
                __syncwarp();
                p += P.cs1 - P.cs0;
            }
            if (prof && blockIdx.x == 0 && leader) {
                g_dense_dbg[8] = (unsigned long long)(clock64() - tstart);
                g_dense_dbg[9] = (unsigned long long)wAcc;
                g_dense_dbg[10] = (unsigned long long)wFull;
                g_dense_dbg[11] = 0;
                g_dense_dbg[12] = (unsigned long long)wIssue;
                g_dense_dbg[14] = it;
            }
        }
        __syncwarp();
    } else if (warp >= 2 && warp < 6) {
        // ===== epilogue: D rows are t * 32 + f, so quadrant q holds term q of all 32 factors.  Per 32 genes the three term
        //       warps park their [factor][gene] blocks in shared memory, the four warps then sum the terms of 8 genes each and
        //       add the row slices into R with one coalesced red.global.add.f32 per gene (R also sums over the CTAs that share
        //       a gene block) =====
        const int q = warp & 3, ew = warp - 2;
        uint32_t icount = 0;
        for (long long p = p0; p < p1; ++icount) {
            const K2Pos P = k2_item(ds, nds, p, p1);
            const DenseDs& D = ds[P.i];
            mbar_wait(accfull, icount & 1u);
            tc_fence_after();
            const int g0 = P.gb * 512;
            const int ngene = min(512, D.rows - g0);
            for (int j = 0; j < 512 / 32; ++j) {
                if (j * 32 >= ngene) break;
                float* buf = smEpi + (j & 1) * (3 * 32 * 33);
                if (q < 3) {
                    uint32_t v[32];
                    tmem_ld32(tmem0 + ((uint32_t)(q * 32) << 16) + j * 32, v);
#pragma unroll
                    for (int c = 0; c < 32; ++c) buf[(q * 32 + lane) * 33 + c] = __uint_as_float(v[c]);
                }
                asm volatile("bar.sync 5, 128;" ::: "memory");
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int c = ew * 8 + cc;
                    const int g = g0 + j * 32 + c;
                    if (j * 32 + c < ngene) {
                        const float sum = buf[lane * 33 + c] + buf[(32 + lane) * 33 + c] + buf[(64 + lane) * 33 + c];
                        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(D.R + (size_t)g * LDF + lane), "f"(sum * __ldg(D.rs + g)) : "memory");
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accempty);
            p += P.cs1 - P.cs0;
        }
    } else {
        // ===== scatter warps =====
        const int sw = warp - 6, slot = sw % NST, gt = (sw / NST) * 32 + lane;
        K2Stages I;
        I.init(ds, nds, p0, p1);
        scatter_role(I, slot, gt, lane, smem_u32(smX + slot * XSTAGE_BYTES), empty + slot, full + slot, dbg);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "n"(512) : "memory");
}

// =====================================================================================================================
// k_gram_tc:  G = H^T H (k x k, fp64 result) on the tensor cores -- the dense Gram of the V / W solves (R/cINMF.R:486,497).
// H rows (fp32) are split into three bf16 terms; per 16 cells ONE operand tile [rows t * 32 + f] x [16 cells] serves as both
// operands:  D[128 x 96] += T[128 x 16] * T[0..95, 16]^T, so D[t 32 + f, t' 32 + f'] = sum_c H_t[c, f] H_t'[c, f'] and
// G[f, f'] is the sum of the nine (t, t') blocks (every bf16 x bf16 product is exact; fp32 accumulation in TMEM over at most
// 512 cells, then fp64).  Work item = 512 cells of one dataset; two 96-column accumulators so that the epilogue of an item
// overlaps the MMAs of the next.  In the gather path the same Gram is phase 2 of k_spmm_xh (fp32 register patches).
// =====================================================================================================================
constexpr int GT_CELLS = 512;                 // cells per work item (accumulated in fp32)
constexpr int GT_KS = 4;                      // k-steps per stage: 64 cells
constexpr int GT_STAGE_BYTES = GT_KS * ATILE_BYTES;   // 16 KB
constexpr int GT_CONV_WARPS = 4 * NST;        // four converter warps per slot, one k-step each
constexpr int GT_WARPS = 5 + GT_CONV_WARPS;   // MMA, 4 epilogue, converters
constexpr size_t GT_SMEM = (size_t)NST * GT_STAGE_BYTES + 32 * 8 + 1024;
constexpr uint32_t IDESC_GRAM = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(96 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

template <int LDF>
__global__ void __launch_bounds__(GT_WARPS * 32, 1) k_gram_tc(const DenseDs* __restrict__ ds_g, int nds, int total_items) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ DenseDs sds[MAX_SDS];
    __shared__ double Gacc[32 * 32];
    unsigned char* raw = align1k(smem_dyn);
    unsigned char* smA = raw;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smA + NST * GT_STAGE_BYTES);
    uint64_t* full = bars;              // NST: four converter warps
    uint64_t* empty = bars + NST;       // NST
    uint64_t* accfull = bars + 2 * NST; // 2
    uint64_t* accempty = accfull + 2;   // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(full + s, GT_CONV_WARPS / NST);
            mbar_init(empty + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(accfull + a, 1);
            mbar_init(accempty + a, 4);
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // rows 96..127 of every operand tile stay zero
        uint4* x4 = reinterpret_cast<uint4*>(smA);
        for (int e = threadIdx.x; e < NST * GT_STAGE_BYTES / 16; e += blockDim.x) x4[e] = make_uint4(0u, 0u, 0u, 0u);
        for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) Gacc[e] = 0.0;
    }
    const DenseDs* ds = cache_ds(ds_g, nds, sds);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *tmem_slot;
    // items blockIdx.x, + gridDim.x, ...: (dataset, 512-cell chunk) via gt_item_off; stages of an item: ceil(cells / 64)
    auto item_info = [&](int item, int& i, int& c0, int& ncell) {
        i = find_ds(ds, nds, item, &DenseDs::gh_blk_off);
        c0 = (item - ds[i].gh_blk_off) * GT_CELLS;
        ncell = min(GT_CELLS, ds[i].n - c0);
    };

    if (warp == 0) {
        {   // ===== MMA issuer (converged warp, elected lane for the tcgen05 instructions) =====
            const bool leader = elect_one();
            uint32_t slot = 0, ph = 0, icount = 0, ready = 0;
            const uint64_t adesc0 = umma_desc(smem_u32(smA), 128 * 16, 128);
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++icount) {
                int i, c0, ncell;
                item_info(item, i, c0, ncell);
                const int nst = (ncell + GT_KS * KSTEP - 1) / (GT_KS * KSTEP);
                const uint32_t acc = icount & 1u, accph = (icount >> 1) & 1u;
                mbar_wait(accempty + acc, accph ^ 1u);
                for (int s = 0; s < nst; ++s) {
                    if (!ready) mbar_wait(full + slot, ph);
                    tc_fence_after();
                    const uint64_t ad = adesc0 + (uint64_t)(slot * (GT_STAGE_BYTES >> 4));
                    if (leader) {
                        umma_bf16(tmem0 + acc * 96, ad, ad, IDESC_GRAM, s != 0);
                        umma_bf16(tmem0 + acc * 96, ad + (uint64_t)(ATILE_BYTES >> 4), ad + (uint64_t)(ATILE_BYTES >> 4), IDESC_GRAM, 1u);
                    }
                    const uint32_t ns = slot + 1 == NST ? 0u : slot + 1, nph = ns == 0 ? ph ^ 1u : ph;
                    ready = mbar_test(full + ns, nph);
                    if (leader) {
                        umma_bf16(tmem0 + acc * 96, ad + (uint64_t)(2 * (ATILE_BYTES >> 4)), ad + (uint64_t)(2 * (ATILE_BYTES >> 4)), IDESC_GRAM, 1u);
                        umma_bf16(tmem0 + acc * 96, ad + (uint64_t)(3 * (ATILE_BYTES >> 4)), ad + (uint64_t)(3 * (ATILE_BYTES >> 4)), IDESC_GRAM, 1u);
                        umma_commit(empty + slot);
                    }
                    slot = ns;
                    ph = nph;
                }
                if (leader) umma_commit(accfull + acc);
                __syncwarp();
            }
        }
        __syncwarp();
    } else if (warp < 5) {
        // ===== epilogue: quadrant q = term t of the rows; lane f sums the three column blocks (terms t') of its row and
        //       adds the 32 partial entries G[f, :] into the CTA's fp64 accumulator =====
        const int q = warp & 3;
        uint32_t icount = 0;
        int cur_ds = -1;
        auto flush = [&]() {   // the four epilogue warps together: Gacc -> global G of dataset cur_ds (fp64 atomics)
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (cur_ds >= 0) {
                double* G = ds[cur_ds].G;
                const int k = ds[cur_ds].kloc;
                constexpr int ldf = LDF;
                for (int e = (warp - 1) * 32 + lane; e < 32 * 32; e += 128) {
                    const int fa = e >> 5, fb = e & 31;
                    if (fa < k && fb < k) atomicAdd(&G[fa * ldf + fb], Gacc[e]);
                    Gacc[e] = 0.0;
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        };
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++icount) {
            int i, c0, ncell;
            item_info(item, i, c0, ncell);
            if (i != cur_ds) {
                flush();
                cur_ds = i;
            }
            const uint32_t acc = icount & 1u, accph = (icount >> 1) & 1u;
            mbar_wait(accfull + acc, accph);
            tc_fence_after();
            if (q < 3) {
                float g[32];
#pragma unroll
                for (int f = 0; f < 32; ++f) g[f] = 0.f;
#pragma unroll
                for (int tb = 0; tb < 3; ++tb) {
                    uint32_t v[32];
                    tmem_ld32(tmem0 + ((uint32_t)(q * 32) << 16) + acc * 96 + tb * 32, v);
#pragma unroll
                    for (int f = 0; f < 32; ++f) g[f] += __uint_as_float(v[f]);
                }
#pragma unroll
                for (int f = 0; f < 32; ++f) atomicAdd(&Gacc[f * 32 + lane], (double)g[f]);   // G is symmetric: entry (f, lane) = (lane, f)
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accempty + acc);
        }
        flush();
    } else {
        // ===== converter warps: slot = cw % NST, k-step = cw / NST; lane f reads H[c, f] of 16 cells one round ahead =====
        const int cw = warp - 5, slot = cw % NST, ks = cw / NST;
        uint4* abase = reinterpret_cast<uint4*>(smA + slot * GT_STAGE_BYTES + ks * ATILE_BYTES);
        float hn[KSTEP], csn = 0.f;
        uint4* img = nullptr;
        int item = (int)blockIdx.x - (int)gridDim.x, s = 0, nst = 0, i = 0, c0 = 0, ncell = 0;
        uint32_t it = 0xFFFFFFFFu, cur_it = 0, nxt_it = 0;
        bool cur_valid = false, nxt_valid = false;
        auto fetch = [&](uint32_t& it_out, bool& valid) {   // loads only
            valid = false;
            for (;;) {
                ++s;
                ++it;
                if (s >= nst) {
                    item += (int)gridDim.x;
                    if (item >= total_items) return;
                    item_info(item, i, c0, ncell);
                    nst = (ncell + GT_KS * KSTEP - 1) / (GT_KS * KSTEP);
                    s = 0;
                }
                if ((int)(it % NST) == slot) break;
            }
            valid = true;
            it_out = it;
            constexpr int ldf = LDF;
            const float* __restrict__ H = ds[i].H + (size_t)c0 * ldf;
            const int cc = s * (GT_KS * KSTEP) + ks * KSTEP;
            csn = (lane < KSTEP && cc + lane < ncell) ? __ldg(ds[i].cs + c0 + cc + lane) : 0.f;
            // image of k_dense_xh's A slabs: [32-cell stage][k-step][chunk of 8 cells][row t * 32 + f][cell % 8], 6 KB per stage
            img = reinterpret_cast<uint4*>(ds[i].Hs) + ((size_t)((c0 + cc) >> 5) * 6144 + (size_t)(((c0 + cc) >> 4) & 1) * 3072) / 16;
#pragma unroll
            for (int j = 0; j < KSTEP; ++j) hn[j] = (cc + j < ncell) ? __ldg(H + (size_t)(cc + j) * ldf + lane) : 0.f;
        };
        fetch(cur_it, cur_valid);
        while (cur_valid) {
            uint4 out[2][3];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                unsigned short hi[8], mid[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split3(hn[ch * 8 + e], hi[e], mid[e], lo[e]);
                out[ch][0] = make_uint4(hi[0] | (uint32_t)hi[1] << 16, hi[2] | (uint32_t)hi[3] << 16, hi[4] | (uint32_t)hi[5] << 16,
                                        hi[6] | (uint32_t)hi[7] << 16);
                out[ch][1] = make_uint4(mid[0] | (uint32_t)mid[1] << 16, mid[2] | (uint32_t)mid[3] << 16, mid[4] | (uint32_t)mid[5] << 16,
                                        mid[6] | (uint32_t)mid[7] << 16);
                out[ch][2] = make_uint4(lo[0] | (uint32_t)lo[1] << 16, lo[2] | (uint32_t)lo[3] << 16, lo[4] | (uint32_t)lo[5] << 16,
                                        lo[6] | (uint32_t)lo[7] << 16);
            }
            mbar_wait(empty + slot, ((cur_it / NST) & 1u) ^ 1u);
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                for (int t = 0; t < 3; ++t) abase[ch * 128 + t * 32 + lane] = out[ch][t];
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full + slot);
            // the same rows scaled by the cell scale, split again and written out: the A slabs of the R sweep (k_dense_xh)
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                unsigned short hi[8], mid[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split3(hn[ch * 8 + e] * __shfl_sync(0xffffffffu, csn, ch * 8 + e), hi[e], mid[e], lo[e]);
                uint4* dst = img + ch * 96;
                dst[lane] = make_uint4(hi[0] | (uint32_t)hi[1] << 16, hi[2] | (uint32_t)hi[3] << 16, hi[4] | (uint32_t)hi[5] << 16,
                                       hi[6] | (uint32_t)hi[7] << 16);
                dst[32 + lane] = make_uint4(mid[0] | (uint32_t)mid[1] << 16, mid[2] | (uint32_t)mid[3] << 16, mid[4] | (uint32_t)mid[5] << 16,
                                            mid[6] | (uint32_t)mid[7] << 16);
                dst[64 + lane] = make_uint4(lo[0] | (uint32_t)lo[1] << 16, lo[2] | (uint32_t)lo[3] << 16, lo[4] | (uint32_t)lo[5] << 16,
                                            lo[6] | (uint32_t)lo[7] << 16);
            }
            fetch(nxt_it, nxt_valid);   // after the publish: the fence must not wait for these loads
            cur_it = nxt_it;
            cur_valid = nxt_valid;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "n"(256) : "memory");
}

// =====================================================================================================================
// k_gram_cross (KP = 64 only):  the off-diagonal block of G = H^T H between the two factor halves, G[f, 32 + f'] and its
// mirror image.  k_gram_tc covers the diagonal blocks (one descriptor per half); the cross block is 2 n 32^2 flops per
// dataset and runs on the CUDA cores: a CTA stages 64 cells x 64 factors at a time, every thread owns a 2 x 2 patch of the
// 32 x 32 block (fp32 over the CTA's cells in chunks of 64, fp64 across chunks), one fp64 atomic per entry and CTA.
// Descriptors 2 i and 2 i + 1 are the halves of dataset i; blockIdx.y = i.  The grid gives a CTA ~32 chunks: with fewer the
// kernel is nothing but its final atomics.
// =====================================================================================================================
constexpr int GX_CHUNKS_PER_CTA = 32;
__global__ void __launch_bounds__(256) k_gram_cross(const DenseDs* __restrict__ ds) {
    const DenseDs& D0 = ds[2 * blockIdx.y];
    const DenseDs& D1 = ds[2 * blockIdx.y + 1];
    const int n = D0.n, k0 = D0.kloc, k1 = D1.kloc;
    if (k1 <= 0) return;
    __shared__ __align__(16) float Hs[64][68];
    const float* __restrict__ H = D0.H;   // row stride 64; factors 0..31 | 32..63
    // 64 threads cover the 32 x 32 block with 4 x 4 patches; the four groups of 64 threads take 16 of the 64 staged cells each
    const int grp = threadIdx.x >> 6, t = threadIdx.x & 63, a0 = (t >> 3) * 4, b0 = (t & 7) * 4;
    double g[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) g[a][b] = 0.0;
    for (int c0 = blockIdx.x * 64; c0 < n; c0 += gridDim.x * 64) {
        __syncthreads();
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            const int r = e >> 4, q = e & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + r < n) v = __ldg(reinterpret_cast<const float4*>(H + (size_t)(c0 + r) * 64) + q);
            *reinterpret_cast<float4*>(&Hs[r][4 * q]) = v;
        }
        __syncthreads();
        float s[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) s[a][b] = 0.f;
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            const int r = grp * 16 + rr;
            const float4 x = *reinterpret_cast<const float4*>(&Hs[r][a0]);
            const float4 y = *reinterpret_cast<const float4*>(&Hs[r][32 + b0]);
            const float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) s[a][b] = fmaf(xs[a], ys[b], s[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) g[a][b] += (double)s[a][b];
    }
    // the four row groups are summed in shared memory (it is free now), then one atomic pair per entry
    __syncthreads();
    double* red = reinterpret_cast<double*>(&Hs[0][0]);   // 64 threads x 16 doubles = 8 KB, one row group at a time
    static_assert(sizeof(Hs) >= 64 * 16 * sizeof(double), "reduction buffer");
    for (int gsrc = 1; gsrc < 4; ++gsrc) {
        if (grp == gsrc) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) red[(a * 4 + b) * 64 + t] = g[a][b];
        }
        __syncthreads();
        if (grp == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) g[a][b] += red[(a * 4 + b) * 64 + t];
        }
        __syncthreads();
    }
    if (grp != 0) return;
    double* G = D0.G;   // 64 x 64
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (a0 + a < k0 && b0 + b < k1) {
                atomicAdd(&G[(a0 + a) * 64 + 32 + b0 + b], g[a][b]);
                atomicAdd(&G[(32 + b0 + b) * 64 + a0 + a], g[a][b]);
            }
}

// =====================================================================================================================
// launches
// =====================================================================================================================
void launch_dense_ltx(const DenseDs* ds, int nds, int total_tiles, int ldf, int num_sms, cudaStream_t st) {
    if (total_tiles <= 0) return;
    const int grid = std::min(total_tiles, num_sms);
    static const int dbg = (getenv("LGR_DENSE_DBG") ? atoi(getenv("LGR_DENSE_DBG")) : 0) |
                           ((getenv("LGR_PF_AHEAD") ? atoi(getenv("LGR_PF_AHEAD")) / 256 : PF_AHEAD / 256) << 8);   // developer knob: 1 no scatter, 2 no MMA, 4 no H loads
    if (ldf == 32)
        k_dense_ltx<32><<<grid, K1_WARPS * 32, K1_SMEM, st>>>(ds, nds, total_tiles, dbg);
    else
        k_dense_ltx<64><<<grid, K1_WARPS * 32, K1_SMEM, st>>>(ds, nds, total_tiles, dbg);
    LGR_CUDA(cudaGetLastError());
    if (dbg & 16) {
        unsigned long long h[16];
        LGR_CUDA(cudaStreamSynchronize(st));
        LGR_CUDA(cudaMemcpyFromSymbol(h, g_dense_dbg, sizeof(h)));
        fprintf(stderr, "[dense dbg] ltx CTA0 MMA thread: total %llu cyc, stages %llu | wait acc %llu, wait full %llu, issue %llu\n", h[0], h[6], h[1],
                h[2], h[4]);
    }
}
void launch_dense_xh(const DenseDs* ds, int nds, long long total_flat, int ldf, int num_sms, cudaStream_t st) {
    if (total_flat <= 0) return;
    const int grid = (int)std::min<long long>(total_flat, num_sms);
    static const int dbg = (getenv("LGR_DENSE_DBG") ? atoi(getenv("LGR_DENSE_DBG")) : 0) |
                           ((getenv("LGR_PF_AHEAD") ? atoi(getenv("LGR_PF_AHEAD")) / 256 : PF_AHEAD / 256) << 8);
    if (ldf == 32)
        k_dense_xh<32><<<grid, K2_WARPS_TOTAL * 32, K2_SMEM, st>>>(ds, nds, total_flat, dbg);
    else
        k_dense_xh<64><<<grid, K2_WARPS_TOTAL * 32, K2_SMEM, st>>>(ds, nds, total_flat, dbg);
    LGR_CUDA(cudaGetLastError());
    if (dbg & 16) {
        unsigned long long h[16];
        LGR_CUDA(cudaStreamSynchronize(st));
        LGR_CUDA(cudaMemcpyFromSymbol(h, g_dense_dbg, sizeof(h)));
        fprintf(stderr, "[dense dbg] xh  CTA0 MMA thread: total %llu cyc, stages %llu | wait acc %llu, wait full %llu, issue %llu\n", h[8], h[14], h[9],
                h[10], h[12]);
    }
}
void launch_gram_h(const DenseDs* ds, int nds, int total_items, int ldf, int num_sms, cudaStream_t st) {
    if (total_items <= 0) return;
    const int grid = std::min(total_items, num_sms);
    if (ldf == 32)
        k_gram_tc<32><<<grid, GT_WARPS * 32, GT_SMEM, st>>>(ds, nds, total_items);
    else
        k_gram_tc<64><<<grid, GT_WARPS * 32, GT_SMEM, st>>>(ds, nds, total_items);
    LGR_CUDA(cudaGetLastError());
}
void launch_gram_cross(const DenseDs* ds, int nds, int max_n, cudaStream_t st) {
    if (nds <= 0 || max_n <= 0) return;
    const int chunks = (max_n + 63) / 64;
    const int ctas = std::max(1, std::min(296, (chunks + GX_CHUNKS_PER_CTA - 1) / GX_CHUNKS_PER_CTA));
    k_gram_cross<<<dim3(ctas, nds / 2), 256, 0, st>>>(ds);
    LGR_CUDA(cudaGetLastError());
}
int dense_gram_tile() { return GT_CELLS; }
void dense_init() {
    LGR_CUDA(cudaFuncSetAttribute(k_dense_ltx<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_SMEM));
    LGR_CUDA(cudaFuncSetAttribute(k_dense_ltx<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_SMEM));
    LGR_CUDA(cudaFuncSetAttribute(k_dense_xh<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K2_SMEM));
    LGR_CUDA(cudaFuncSetAttribute(k_dense_xh<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K2_SMEM));
    LGR_CUDA(cudaFuncSetAttribute(k_gram_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
    LGR_CUDA(cudaFuncSetAttribute(k_gram_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
}

// =====================================================================================================================
// one-time layout build (upload path): CSC + scales -> the two entry streams
//   ltx stream: segments (256-cell tile, 64-gene stage); entry offset addresses [k-step][chunk of 8 genes][cell][gene % 8]
//   xh  stream: segments (512-gene block, 32-cell stage); entry offset addresses [k-step][half block][chunk of 8 cells][gene][cell % 8]
// Every entry is verified to be an integer count in 1..256 (else *bad is raised: the scales do not factor X).
// =====================================================================================================================
// Order inside a segment: (rank of the entry among those of its shared-memory bank, bank).  A warp of a scatter group
// stores 32 consecutive entries at a time, so consecutive entries must fall into different banks: with this order the
// first rounds of a segment touch every bank once, and only the tail (the fullest banks) repeats banks.
template <typename KeyT>
__global__ void k_dense_keys(int n, int rows, const int* __restrict__ colptr, const int* __restrict__ rowidx, const float* __restrict__ val,
                             const float* __restrict__ rs, const float* __restrict__ cs, int nstage1, int ncstage,
                             KeyT* __restrict__ key1, unsigned int* __restrict__ ent1, int* __restrict__ cnt1, int* __restrict__ bank1,
                             KeyT* __restrict__ key2, unsigned int* __restrict__ ent2, int* __restrict__ cnt2, int* __restrict__ bank2,
                             int* __restrict__ bad) {
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
            const unsigned int bank = (off >> 1) & 31u;
            const unsigned int rank = min((unsigned)atomicAdd(&bank1[(size_t)key * 32 + bank], 1), 31u);
            key1[t] = ((KeyT)key << 10) | (KeyT)((rank << 5) | bank);
            ent1[t] = off | (bf << 16);
            atomicAdd(&cnt1[key], 1);
        }
        {   // xh: columns = genes of the half block, k = cells
            const int gl = g % 512, nn = gl % TILE_N, sub = gl / TILE_N, kk = c % (K2_KS * KSTEP);
            const unsigned int off =
                (unsigned)(((kk / KSTEP) * 2 + sub) * XTILE_BYTES + ((kk % KSTEP) / 8) * (TILE_N * 16) + nn * 16 + (kk % 8) * 2) >> 1;
            const unsigned int key = (unsigned)(g / 512) * (unsigned)ncstage + (unsigned)(c / (K2_KS * KSTEP));
            const unsigned int bank = (off >> 1) & 31u;
            const unsigned int rank = min((unsigned)atomicAdd(&bank2[(size_t)key * 32 + bank], 1), 31u);
            key2[t] = ((KeyT)key << 10) | (KeyT)((rank << 5) | bank);
            ent2[t] = off | (bf << 16);
            atomicAdd(&cnt2[key], 1);
        }
    }
}

template <typename KeyT>
static void sort_stream(DBuf<KeyT>& key, DBuf<unsigned int>& ent, DBuf<int>& cnt, size_t nnz, size_t nseg, int key_bits, DBuf<int>& seg,
                        DBuf<unsigned int>& out, cudaStream_t st) {
    DBuf<KeyT> key2(std::max<size_t>(nnz, 1));
    out.alloc(std::max<size_t>(nnz, 1) + (size_t)65536 + 1024);   // slack: the scatter warps prefetch past the end
    seg.alloc(nseg + 1);
    size_t tmp_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, key2.p, ent.p, out.p, (int)nnz, 0, key_bits, st);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, cnt.p, seg.p, (int)(nseg + 1), st);
    DBuf<unsigned char> tmp(std::max(tmp_bytes, scan_bytes));
    LGR_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key.p, key2.p, ent.p, out.p, (int)nnz, 0, key_bits, st));
    LGR_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, cnt.p, seg.p, (int)(nseg + 1), st));
    LGR_CUDA(cudaStreamSynchronize(st));   // the temporaries are released on return
}

template <typename KeyT>
static void dense_build_t(int n, int rows, const int* colptr, const int* rowidx, const float* val, size_t nnz, const float* rs,
                          const float* cs, DenseLayout& out, int* bad, size_t nseg1, size_t nseg2, int bits1, int bits2, cudaStream_t st) {
    const size_t nz = std::max<size_t>(nnz, 1);
    DBuf<KeyT> key1(nz), key2(nz);
    DBuf<unsigned int> ent1(nz), ent2(nz);
    DBuf<int> cnt1(nseg1 + 1), cnt2(nseg2 + 1), bank1(nseg1 * 32), bank2(nseg2 * 32);
    cnt1.zero(st);
    cnt2.zero(st);
    bank1.zero(st);
    bank2.zero(st);
    if (nnz && n > 0) {
        k_dense_keys<KeyT><<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, st>>>(n, rows, colptr, rowidx, val, rs, cs, out.nstage1, out.ncstage,
                                                                                       key1.p, ent1.p, cnt1.p, bank1.p, key2.p, ent2.p, cnt2.p,
                                                                                       bank2.p, bad);
        LGR_CUDA(cudaGetLastError());
    }
    bank1.release();
    bank2.release();
    sort_stream<KeyT>(key1, ent1, cnt1, nnz, nseg1, bits1, out.seg1, out.ent1, st);
    key1.release();
    ent1.release();
    sort_stream<KeyT>(key2, ent2, cnt2, nnz, nseg2, bits2, out.seg2, out.ent2, st);
}

void dense_build(int n, int rows, const int* colptr, const int* rowidx, const float* val, size_t nnz, const float* rs, const float* cs,
                 int halves, DenseLayout& out, int* bad, cudaStream_t st) {
    out.nstage1 = (rows + K1_KS * KSTEP - 1) / (K1_KS * KSTEP);
    out.ntile1 = (n + TILE_N - 1) / TILE_N;
    out.ngb = (rows + 511) / 512;
    out.ncstage = (n + K2_KS * KSTEP - 1) / (K2_KS * KSTEP);
    const size_t nseg1 = (size_t)out.ntile1 * out.nstage1, nseg2 = (size_t)out.ngb * out.ncstage;
    LGR_REQUIRE(nseg1 < (size_t)0x7fffffff / 32 && nseg2 < (size_t)0x7fffffff / 32 && nnz < (size_t)0x7fffffff, "dense layout: too many segments");
    auto bits_of = [](size_t nseg) {
        int b = 1;
        while (((size_t)1 << b) < nseg) ++b;
        return b + 10;   // segment | rank (5 bits, clamped: the tail of an over-full bank may collide) | bank (5 bits)
    };
    const int bits1 = bits_of(nseg1), bits2 = bits_of(nseg2);
    if (bits1 <= 32 && bits2 <= 32)
        dense_build_t<unsigned int>(n, rows, colptr, rowidx, val, nnz, rs, cs, out, bad, nseg1, nseg2, bits1, bits2, st);
    else
        dense_build_t<unsigned long long>(n, rows, colptr, rowidx, val, nnz, rs, cs, out, bad, nseg1, nseg2, bits1, bits2, st);
    // pre-split factor image of k_dense_ltx: nstage1 x 16 KB, zero (rows of term 3 and rows >= `rows` stay zero)
    // (KP = 64: one image per half of the factors, back to back)
    out.Asplit.alloc((size_t)halves * out.nstage1 * K1_ASTAGE_BYTES / 2);
    out.Asplit.zero(st);
    // split image of cs * H for k_dense_xh (written by k_gram_tc every iteration): 6 KB per 32-cell stage; k_gram_tc covers
    // whole 64-cell stages, so the last stage pair is allocated in full
    out.Hs.alloc((size_t)halves * ((size_t)out.ncstage + 2) * 3072);
    out.Hs.zero(st);
}

}  // namespace lgr

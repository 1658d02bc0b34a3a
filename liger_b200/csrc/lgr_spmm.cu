// liger_b200 -- the two sparse contractions of an ANLS iteration (sm_100a), both as shared-memory GATHERS:
//
//   k_spmm_ltx_tiled   B = L~^T X   (rhs of the H solve, reference R/cINMF.R:480)
//       X in "gene-tiled CSC": genes are cut into tiles of GT rows whose L~ rows (GT x KP fp32) fit in shared
//       memory; inside a gene tile the entries are ordered by cell.  One CTA owns a block of CB cells, walks the
//       gene tiles, stages each L~ tile with one TMA bulk copy and accumulates the partial B rows in shared memory.
//   k_spmm_xh          R = X H  and  G = H^T H   (statistics of the V / W solves, R/cINMF.R:486-488,497-499)
//       X in "cell-tiled CSR": cells are cut into tiles of TC whose H rows (TC x KP fp32, contiguous in HBM) are
//       staged by one TMA bulk copy; inside a cell tile the entries are ordered by gene.  Warps own genes, gather
//       H rows, and issue one vector RED (red.global.add.v4.f32) per (gene, tile).
//
// Both formats pad every (tile, minor) segment to a multiple of 8 entries (u16 local index + fp32 value, zero padded)
// so that a quarter-warp fetches its 8 entries with three aligned 128-bit loads and then issues, per non-zero,
// one 128-bit shared-memory load and four FFMAs: ~2 warp instructions and one shared-memory wavefront per non-zero.
#include "lgr_common.cuh"
#include "lgr_ptx.cuh"

namespace lgr {

// acc += sum_j val_j * tile[loc_j / LPN][4s..4s+3]   for the 8 entries of one group.
// Local indices are stored PRE-SCALED by LPN (= float4 slots per staged row) and `tile4s` already points at this
// lane's 16-byte slot, so the shared-memory address of a non-zero is one mask/shift plus one LEA.
__device__ __forceinline__ void gather8(float4& acc, const uint4 cells, const float4 v0, const float4 v1,
                                        const float4* __restrict__ tile4s) {
    const unsigned cw[4] = {cells.x, cells.y, cells.z, cells.w};
    const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const unsigned c = (j & 1) ? (cw[j >> 1] >> 16) : (cw[j >> 1] & 0xffffu);
        const float4 h = tile4s[c];
        acc.x = fmaf(vv[j], h.x, acc.x);
        acc.y = fmaf(vv[j], h.y, acc.y);
        acc.z = fmaf(vv[j], h.z, acc.z);
        acc.w = fmaf(vv[j], h.w, acc.w);
    }
}

struct Group8 {
    uint4 cells;
    float4 v0, v1;
};
__device__ __forceinline__ Group8 load_group(const unsigned short* __restrict__ loc, const float* __restrict__ val, int g) {
    Group8 r;
    r.cells = __ldg(reinterpret_cast<const uint4*>(loc + (size_t)g * 8));
    r.v0 = __ldg(reinterpret_cast<const float4*>(val + (size_t)g * 8));
    r.v1 = __ldg(reinterpret_cast<const float4*>(val + (size_t)g * 8 + 4));
    return r;
}

// A lane-group (KP/4 lanes) walks the contiguous stream of 8-entry groups of the segments [sa, sb): every group is
// one 128-bit load of local indices, two of values (prefetched one group ahead, two register buffers in ping-pong),
// then per non-zero one 128-bit shared-memory load of the staged factor row and four FFMAs.  At a segment end the
// 32 (64) sums the lane-group holds are the complete output row, so there is no cross-lane reduction;
// `store(seg, acc, nonempty)` consumes them.
template <int KP, typename Store>
__device__ __forceinline__ void gather_stream(const int* __restrict__ ptr8, const int sa, const int sb,
                                              const unsigned short* __restrict__ loc, const float* __restrict__ val,
                                              const float4* __restrict__ tile4, const int s, Store store) {
    const float4* __restrict__ tile4s = tile4 + s;
    int seg = sa;
    int g = 0, gend = 0, segend = 0, nsegend = 0;
    if (sa < sb) {
        g = __ldg(ptr8 + sa);
        gend = __ldg(ptr8 + sb);
        segend = __ldg(ptr8 + sa + 1);
        nsegend = (sa + 2 <= sb) ? __ldg(ptr8 + sa + 2) : gend;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    bool nonempty = false;
    const int glast = gend > 0 ? gend - 1 : 0;   // prefetches are clamped, never predicated
    bool alive = g < gend;
    if (!alive)
        for (; seg < sb; ++seg) store(seg, acc, false);   // nothing but empty segments
    Group8 a = load_group(loc, val, alive ? g : 0), b;
    // One group per step and per live lane-group.  The warp-wide vote at the top of every step is the reconvergence
    // point: lane-groups whose segments end at different places fall back into lockstep immediately.
#define LGR_GATHER_STEP(CUR, NXT)                                                          \
    if (alive) {                                                                           \
        while (g >= segend) { /* close the segments that end before this group */          \
            store(seg, acc, nonempty);                                                     \
            acc = make_float4(0.f, 0.f, 0.f, 0.f);                                         \
            nonempty = false;                                                              \
            ++seg;                                                                         \
            segend = nsegend;                                                              \
            nsegend = (seg + 2 <= sb) ? __ldg(ptr8 + seg + 2) : gend;                      \
        }                                                                                  \
        NXT = load_group(loc, val, min(g + 1, glast));                                     \
        gather8(acc, CUR.cells, CUR.v0, CUR.v1, tile4s);                                   \
        nonempty = true;                                                                   \
        if (++g >= gend) { /* stream exhausted: close the current and the trailing empty segments */ \
            for (; seg < sb; ++seg) {                                                      \
                store(seg, acc, nonempty);                                                 \
                acc = make_float4(0.f, 0.f, 0.f, 0.f);                                     \
                nonempty = false;                                                          \
            }                                                                              \
            alive = false;                                                                 \
        }                                                                                  \
    }
    while (__any_sync(0xffffffffu, alive)) {
        LGR_GATHER_STEP(a, b)
        if (!__any_sync(0xffffffffu, alive)) break;
        LGR_GATHER_STEP(b, a)
    }
#undef LGR_GATHER_STEP
}

// ------------------------------------------------------------------------------------------------
// k_spmm_ltx_tiled
// ------------------------------------------------------------------------------------------------
constexpr int LTX_THREADS = 1024;

size_t spmm_ltx_smem(int kp, int cb, int gt) { return (size_t)(cb + gt) * kp * sizeof(float) + 64; }

template <int KP>
__global__ void __launch_bounds__(LTX_THREADS, 1) k_spmm_ltx_tiled(const DsDev* __restrict__ ds, int nds, int CB, int GT) {
    constexpr int LPN = KP / 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Ls = reinterpret_cast<float*>(smem_raw);             // GT x KP
    float* Bp = Ls + (size_t)GT * KP;                            // CB x KP partial rows of B
    uint64_t* bar = reinterpret_cast<uint64_t*>(Bp + (size_t)CB * KP);
    const int i = find_ds(ds, nds, (int)blockIdx.x, &DsDev::ltx_blk_off);
    const DsDev& D = ds[i];
    const int n = D.n, rows = D.rows, ngt = D.ngt;
    const int c0 = ((int)blockIdx.x - D.ltx_blk_off) * CB;
    const int ncell = min(CB, n - c0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int q = lane / LPN, s = lane % LPN;
    const float4* Ls4 = reinterpret_cast<const float4*>(Ls);
    float4* Bp4 = reinterpret_cast<float4*>(Bp);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    // each warp owns a contiguous run of cells of the block (contiguous entries in HBM)
    const int cpw = (ncell + nwarps - 1) / nwarps;
    const int w0 = min(warp * cpw, ncell), w1 = min(w0 + cpw, ncell);
    for (int gt = 0; gt < ngt; ++gt) {
        __syncthreads();  // everyone is done reading the previous L~ tile (and sees the barrier init)
        if (threadIdx.x == 0) {
            const int r0 = gt * GT;
            const uint32_t bytes = (uint32_t)min(GT, rows - r0) * KP * sizeof(float);
            mbar_expect_tx(bar, bytes);
            const char* src = reinterpret_cast<const char*>(D.L + (size_t)r0 * KP);
            char* dst = reinterpret_cast<char*>(Ls);
            for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(dst + off, src + off, min(32768u, bytes - off), bar);
        }
        const int* __restrict__ ptr = D.cptr8 + (size_t)gt * n + c0;
        mbar_wait(bar, (uint32_t)(gt & 1));
        {
            // the warp's cells [w0, w1) are split into contiguous runs, one per lane-group
            const int nq = 32 / LPN;
            const int per = (w1 - w0 + nq - 1) / nq;
            const int sa = min(w0 + q * per, w1), sb = min(sa + per, w1);
            const bool first = gt == 0;
            gather_stream<KP>(ptr, sa, sb, D.crow, D.cval, Ls4, s, [&](int cell, float4 acc, bool) {
                float4* dst = Bp4 + (size_t)cell * LPN + s;
                if (!first) {
                    const float4 o = *dst;
                    acc.x += o.x;
                    acc.y += o.y;
                    acc.z += o.z;
                    acc.w += o.w;
                }
                *dst = acc;
            });
        }
    }
    __syncthreads();
    float4* __restrict__ out = reinterpret_cast<float4*>(D.B) + (size_t)c0 * LPN;
    for (int e = threadIdx.x; e < ncell * LPN; e += blockDim.x) out[e] = Bp4[e];
}

void launch_spmm_ltx_tiled(const DsDev* ds, int nds, int total_blocks, int kp, int cb, int gt, cudaStream_t st) {
    if (total_blocks <= 0) return;
    const size_t smem = spmm_ltx_smem(kp, cb, gt);
    if (kp == 32)
        k_spmm_ltx_tiled<32><<<total_blocks, LTX_THREADS, smem, st>>>(ds, nds, cb, gt);
    else
        k_spmm_ltx_tiled<64><<<total_blocks, LTX_THREADS, smem, st>>>(ds, nds, cb, gt);
    LGR_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// k_spmm_xh
// ------------------------------------------------------------------------------------------------
constexpr int XH_THREADS = 1024;
constexpr int XH_MAX_CG = 14;  // cell groups of the H^T H phase (bounds its reduction scratch)

static int xh_ncg(int kp) {
    const int np = kp / 4, npatch = np * (np + 1) / 2;
    int ncg = XH_THREADS / npatch;
    if (ncg < 1) ncg = 1;
    if (ncg > XH_MAX_CG) ncg = XH_MAX_CG;
    return ncg;
}
size_t spmm_xh_smem(int kp, int tc) {
    const int np = kp / 4, npatch = np * (np + 1) / 2;
    return (size_t)tc * kp * sizeof(float) + (size_t)xh_ncg(kp) * npatch * 16 * sizeof(float) + 64;
}

template <int KP>
__global__ void __launch_bounds__(XH_THREADS, 1) k_spmm_xh(const DsDev* __restrict__ ds, int nds, int k, int TC, int ncg) {
    constexpr int LPN = KP / 4;
    constexpr int NP = KP / 4;                 // 4-wide patches per side
    constexpr int NPATCH = NP * (NP + 1) / 2;  // upper-triangular patches
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Hs = reinterpret_cast<float*>(smem_raw);
    float* red = Hs + (size_t)TC * KP;
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + (size_t)ncg * NPATCH * 16);
    __shared__ int next_blk;

    const int i = find_ds(ds, nds, (int)blockIdx.x, &DsDev::xh_blk_off);
    const DsDev& D = ds[i];
    const int tile = (int)blockIdx.x - D.xh_blk_off;
    const int c0 = tile * TC;
    const int ncell = min(TC, D.n - c0);
    const int rows = D.rows;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        next_blk = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)ncell * KP * sizeof(float);
        mbar_expect_tx(bar, bytes);
        const char* src = reinterpret_cast<const char*>(D.H + (size_t)c0 * KP);
        char* dst = reinterpret_cast<char*>(Hs);
        for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(dst + off, src + off, min(32768u, bytes - off), bar);
    }
    mbar_wait(bar, 0);

    // ---- phase 1: R += X_tile * Hs ; warps pull blocks of 32 genes ----
    {
        const int q = lane / LPN, s = lane % LPN;
        const int* __restrict__ tptr = D.tptr8 + (size_t)tile * rows;
        const float4* Hs4 = reinterpret_cast<const float4*>(Hs);
        const int nblk = (rows + 31) >> 5;
        for (;;) {
            int b = 0;
            if (lane == 0) b = atomicAdd(&next_blk, 1);
            b = __shfl_sync(0xffffffffu, b, 0);
            if (b >= nblk) break;
            const int g0 = b * 32;
            const int ng = min(32, rows - g0);
            const int nq = 32 / LPN;
            const int per = (ng + nq - 1) / nq;
            const int sa = min(g0 + q * per, g0 + ng), sb = min(sa + per, g0 + ng);
            float* __restrict__ R = D.R;
            gather_stream<KP>(tptr, sa, sb, D.tcell, D.tval, Hs4, s, [&](int gene, float4 acc, bool nonempty) {
                if (nonempty) red_add_v4(R + (size_t)gene * KP + s * 4, acc);
            });
        }
    }

    // ---- phase 2: G += Hs^T Hs over this tile (4x4 register patches of the upper triangle) ----
    {
        const int t = threadIdx.x;
        const int cg = t / NPATCH, pt = t % NPATCH;
        int pi = 0, pj = 0;
        {
            int rem = pt;
            while (rem >= NP - pi) {
                rem -= NP - pi;
                ++pi;
            }
            pj = pi + rem;
        }
        const float4* Hs4 = reinterpret_cast<const float4*>(Hs);
        if (cg < ncg) {
            float a[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) a[x][y] = 0.f;
            for (int c = cg; c < ncell; c += ncg) {
                const float4 u = Hs4[c * NP + pi];
                const float4 v = Hs4[c * NP + pj];
                const float uu[4] = {u.x, u.y, u.z, u.w};
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) a[x][y] = fmaf(uu[x], vv[y], a[x][y]);
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) red[((size_t)cg * NPATCH + pt) * 16 + x * 4 + y] = a[x][y];
        }
        __syncthreads();
        for (int e = t; e < NPATCH * 16; e += blockDim.x) {
            const int p = e / 16, xy = e % 16, x = xy / 4, y = xy % 4;
            double sum = 0.0;
            for (int g = 0; g < ncg; ++g) sum += (double)red[((size_t)g * NPATCH + p) * 16 + xy];
            int ppi = 0, rem = p;
            while (rem >= NP - ppi) {
                rem -= NP - ppi;
                ++ppi;
            }
            const int ppj = ppi + rem;
            const int fa = ppi * 4 + x, fb = ppj * 4 + y;
            if (fa < k && fb < k) {
                atomicAdd(&D.G[fa * KP + fb], sum);
                if (ppi != ppj) atomicAdd(&D.G[fb * KP + fa], sum);  // a diagonal patch already holds both triangles
            }
        }
    }
}

void launch_spmm_xh(const DsDev* ds, int nds, int total_tiles, int k, int kp, int tc, cudaStream_t st) {
    if (total_tiles <= 0) return;
    const size_t smem = spmm_xh_smem(kp, tc);
    const int ncg = xh_ncg(kp);
    if (kp == 32)
        k_spmm_xh<32><<<total_tiles, XH_THREADS, smem, st>>>(ds, nds, k, tc, ncg);
    else
        k_spmm_xh<64><<<total_tiles, XH_THREADS, smem, st>>>(ds, nds, k, tc, ncg);
    LGR_CUDA(cudaGetLastError());
}

void spmm_init() {
    const int maxs = 227 * 1024 - 1024;
    LGR_CUDA(cudaFuncSetAttribute(k_spmm_xh<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    LGR_CUDA(cudaFuncSetAttribute(k_spmm_xh<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    LGR_CUDA(cudaFuncSetAttribute(k_spmm_ltx_tiled<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    LGR_CUDA(cudaFuncSetAttribute(k_spmm_ltx_tiled<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
}

}  // namespace lgr

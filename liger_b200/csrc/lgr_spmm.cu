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
// Both formats pad every (tile, minor) segment to a multiple of 8 entries (16-bit pre-scaled local index + fp32 value,
// zero padded) and keep one 32-bit tag per group of 8 = the group's output row (lgr_layout.cu).  The stream of a CTA is
// therefore a flat, self-describing list of groups that is cut into EQUAL contiguous ranges, one per lane-group
// (KP/4 lanes = one staged row per 128-bit shared-memory load): uniform trip counts, no votes, no pointer walks.
// Per group: three 128-bit loads + the tag (prefetched one group ahead), then per non-zero ~2 ALU ops, one 128-bit shared-memory
// load and four FFMAs; a change of the tag flushes the lane-group's 4 sums (the complete output row slice).
#include "lgr_common.cuh"
#include "lgr_ptx.cuh"

namespace lgr {

struct Group8 {
    uint4 i;        // 8 x 16-bit pre-scaled local indices
    float4 v0, v1;  // 8 values
    unsigned tag;   // output row of the group
};
__device__ __forceinline__ Group8 load_group(const unsigned short* __restrict__ loc, const unsigned int* __restrict__ tag,
                                             const float* __restrict__ val, int g) {
    Group8 r;
    const float4* vi = reinterpret_cast<const float4*>(val + (size_t)g * 8);
    r.i = __ldg(reinterpret_cast<const uint4*>(loc + (size_t)g * 8));
    r.v0 = __ldg(vi);
    r.v1 = __ldg(vi + 1);
    r.tag = __ldg(tag + g);
    return r;
}

// acc += sum_j val_j * tile[idx_j][4s..4s+3]; `tile4s` already points at this lane's 16-byte slot of row 0.
__device__ __forceinline__ void gather8(float4& acc, const Group8& G, const float4* __restrict__ tile4s) {
    const unsigned c[8] = {G.i.x & 0xffffu, G.i.x >> 16, G.i.y & 0xffffu, G.i.y >> 16,
                           G.i.z & 0xffffu, G.i.z >> 16, G.i.w & 0xffffu, G.i.w >> 16};
    const float vv[8] = {G.v0.x, G.v0.y, G.v0.z, G.v0.w, G.v1.x, G.v1.y, G.v1.z, G.v1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 h = tile4s[c[j]];
        acc.x = fmaf(vv[j], h.x, acc.x);
        acc.y = fmaf(vv[j], h.y, acc.y);
        acc.z = fmaf(vv[j], h.z, acc.z);
        acc.w = fmaf(vv[j], h.w, acc.w);
    }
}

// One lane-group consumes the groups [g, gend) of the stream; `trips` is the warp-uniform loop count (>= gend - g)
// and `glast` the last group the prefetch may touch.  flush(row, acc, edge): `edge` is set for the first and the
// last flush of the range, whose segments may continue in a neighbouring range.
template <typename Flush>
__device__ __forceinline__ void gather_range(const unsigned short* __restrict__ loc, const unsigned int* __restrict__ tag,
                                             const float* __restrict__ val, int g, const int gend, const int trips, const int glast,
                                             const float4* __restrict__ tile4s, Flush flush) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool any = g < gend;
    Group8 cur = load_group(loc, tag, val, min(g, glast));
    unsigned seg = cur.tag;
    bool first = true;
#pragma unroll 2
    for (int i = 0; i < trips; ++i) {
        const Group8 nxt = load_group(loc, tag, val, min(g + 1, glast));
        if (g < gend) {
            const unsigned t = cur.tag;
            if (t != seg) {
                flush(seg, acc, first);
                first = false;
                acc = make_float4(0.f, 0.f, 0.f, 0.f);
                seg = t;
            }
            gather8(acc, cur, tile4s);
        }
        cur = nxt;
        ++g;
    }
    if (any) flush(seg, acc, true);
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
    for (int e = threadIdx.x; e < ncell * LPN; e += blockDim.x) Bp4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nq = 32 / LPN, NQ = nwarps * nq, qid = warp * nq + q;
    for (int gt = 0; gt < ngt; ++gt) {
        __syncthreads();  // everyone is done reading the previous L~ tile (and sees the barrier init / the zeroed Bp)
        if (threadIdx.x == 0) {
            const int r0 = gt * GT;
            const uint32_t bytes = (uint32_t)min(GT, rows - r0) * KP * sizeof(float);
            mbar_expect_tx(bar, bytes);
            const char* src = reinterpret_cast<const char*>(D.L + (size_t)r0 * KP);
            char* dst = reinterpret_cast<char*>(Ls);
            for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(dst + off, src + off, min(32768u, bytes - off), bar);
        }
        // this (gene tile, cell block)'s groups, cut into NQ equal ranges
        const int* __restrict__ ptr = D.cptr8 + (size_t)gt * n + c0;
        const int G0 = __ldg(ptr), G1 = __ldg(ptr + ncell);
        const int per = (G1 - G0 + NQ - 1) / NQ;
        const int g = min(G0 + qid * per, G1), gend = min(g + per, G1);
        mbar_wait(bar, (uint32_t)(gt & 1));
        gather_range(D.crow, D.ctag, D.cval, g, gend, per, max(G1 - 1, 0), Ls4 + s, [&](unsigned cell, float4 acc, bool edge) {
            float* dst = Bp + (size_t)cell * KP + s * 4;
            if (edge) {   // the segment may be shared with the neighbouring range
                atomicAdd(dst + 0, acc.x);
                atomicAdd(dst + 1, acc.y);
                atomicAdd(dst + 2, acc.z);
                atomicAdd(dst + 3, acc.w);
            } else {
                float4* d4 = reinterpret_cast<float4*>(dst);
                float4 o = *d4;
                o.x += acc.x;
                o.y += acc.y;
                o.z += acc.z;
                o.w += acc.w;
                *d4 = o;
            }
        });
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

    const int i = find_ds(ds, nds, (int)blockIdx.x, &DsDev::xh_blk_off);
    const DsDev& D = ds[i];
    const int tile = (int)blockIdx.x - D.xh_blk_off;
    const int c0 = tile * TC;
    const int ncell = min(TC, D.n - c0);
    const int rows = D.rows;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)ncell * KP * sizeof(float);
        mbar_expect_tx(bar, bytes);
        const char* src = reinterpret_cast<const char*>(D.H + (size_t)c0 * KP);
        char* dst = reinterpret_cast<char*>(Hs);
        for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(dst + off, src + off, min(32768u, bytes - off), bar);
    }
    mbar_wait(bar, 0);

    // ---- phase 1: R += X_tile * Hs ; the tile's groups are cut into equal ranges, one per lane-group ----
    {
        const int q = lane / LPN, s = lane % LPN;
        const int nq = 32 / LPN, NQ = (int)(blockDim.x >> 5) * nq, qid = (int)(threadIdx.x >> 5) * nq + q;
        const int* __restrict__ tptr = D.tptr8 + (size_t)tile * rows;
        const int G0 = __ldg(tptr), G1 = __ldg(tptr + rows);
        const int per = (G1 - G0 + NQ - 1) / NQ;
        const int g = min(G0 + qid * per, G1), gend = min(g + per, G1);
        const float4* Hs4 = reinterpret_cast<const float4*>(Hs);
        float* __restrict__ R = D.R;
        gather_range(D.tcell, D.ttag, D.tval, g, gend, per, max(G1 - 1, 0), Hs4 + s,
                     [&](unsigned gene, float4 acc, bool) { red_add_v4(R + (size_t)gene * KP + s * 4, acc); });
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

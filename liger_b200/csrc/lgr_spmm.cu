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

// packed fp32 x2 FMA (Blackwell FFMA2 with a scalar multiplier)
__device__ __forceinline__ void ffma2_s(float& d0, float& d1, float a0, float a1, float b) {
    unsigned long long d, a, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

// acc += sum_j val_j * tile[idx_j][4s..4s+3]; `tile4s` already points at this lane's 16-byte slot of row 0.
__device__ __forceinline__ void gather8(float4& acc, const Group8& G, const float4* __restrict__ tile4s) {
    const unsigned c[8] = {G.i.x & 0xffffu, G.i.x >> 16, G.i.y & 0xffffu, G.i.y >> 16,
                           G.i.z & 0xffffu, G.i.z >> 16, G.i.w & 0xffffu, G.i.w >> 16};
    const float vv[8] = {G.v0.x, G.v0.y, G.v0.z, G.v0.w, G.v1.x, G.v1.y, G.v1.z, G.v1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 h = tile4s[c[j]];
        ffma2_s(acc.x, acc.y, h.x, h.y, vv[j]);
        ffma2_s(acc.z, acc.w, h.z, h.w, vv[j]);
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
// tcgen05 helpers for the fused epilogue  U = B P  (3 x TF32, see lgr_tc.cu for the stand-alone version)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start [0,14) >>4, LBO [16,30) >>4, SBO [32,46) >>4, version [46,48) = 1, no swizzle
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void split_tf32_4(const float4 v, float4& h, float4& l) {
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);   // sign, exponent, 10 leading mantissa bits
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);   // exact in fp32
}

// ------------------------------------------------------------------------------------------------
// k_spmm_ltx_tiled
// ------------------------------------------------------------------------------------------------
constexpr int LTX_THREADS = 1024;

// [L~ tile | B tile | barriers]; with the fused U = B P epilogue the L~ region doubles as the two operand tiles
// (hi / lo parts of B, 2 x CB rows) and the hi / lo parts of P follow the B tile
size_t spmm_ltx_smem(int kp, int cb, int gt, bool fuse_u) {
    const size_t lrows = fuse_u ? (size_t)std::max(gt, 2 * cb) : (size_t)gt;
    return (lrows + cb) * kp * sizeof(float) + (fuse_u ? (size_t)2 * kp * kp * sizeof(float) : 0) + 64;
}

template <int KP, bool FUSE_U>
__global__ void __launch_bounds__(LTX_THREADS, 1) k_spmm_ltx_tiled(const DsDev* __restrict__ ds, int nds, int CB, int GT) {
    constexpr int LPN = KP / 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int LROWS = FUSE_U ? max(GT, 2 * CB) : GT;
    float* Ls = reinterpret_cast<float*>(smem_raw);             // GT x KP (fused epilogue: 2 x CB x KP operand tiles)
    float* Bp = Ls + (size_t)LROWS * KP;                         // CB x KP partial rows of B
    float* Pt = Bp + (size_t)CB * KP;                            // fused epilogue: hi / lo parts of P, KP x KP each
    uint64_t* bar = reinterpret_cast<uint64_t*>(Pt + (FUSE_U ? 2 * KP * KP : 0));
    __shared__ uint32_t tmem_base_s;
    if (FUSE_U && threadIdx.x < 32) {   // 128 TMEM columns: one KP-column accumulator per 128-cell tile
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(128)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int i = find_ds(ds, nds, (int)blockIdx.x, &DsDev::ltx_blk_off);
    const DsDev& D = ds[i];
    const int n = D.n, rows = D.rows, ngt = D.ngt;
    const int c0 = ((int)blockIdx.x - D.ltx_blk_off) * CB;
    const int ncell = min(CB, n - c0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int q = lane / LPN, s = lane % LPN;
    const float4* Ls4 = reinterpret_cast<const float4*>(Ls);
    float4* Bp4 = reinterpret_cast<float4*>(Bp);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        if (FUSE_U) mbar_init(bar + 1, 1);
    }
    for (int e = threadIdx.x; e < (FUSE_U ? CB : ncell) * LPN; e += blockDim.x) Bp4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
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
    if (!FUSE_U) return;

    // ---- fused epilogue (LGR_FLAG_TENSOR_U): U = B P for this block's cells on the tensor cores (u = P b is what the
    //      dual side of the NNLS needs for every column; P = CtC^-1 from k_gram_inv, which runs before this kernel).
    //      3 x TF32: U = B_hi P_hi + B_lo P_hi + B_hi P_lo, fp32 accumulators in TMEM.  Operands in the canonical
    //      K-major no-swizzle UMMA layout: 16-byte chunk (row r, k-chunk c) of a 128-row tile at (c * 128 + r) * 16. ----
    constexpr int M = 128, CH = KP / 4;
    const int ntile = CB / M;
    float4* Ahi = reinterpret_cast<float4*>(Ls);
    float4* Alo = Ahi + (size_t)CB * CH;
    float4* Phi = reinterpret_cast<float4*>(Pt);
    float4* Plo = Phi + KP * CH;
    // B tile (row major) -> operand tiles.  Lanes walk a diagonal of an 8-row x 8-chunk block: the eight reads of a
    // phase hit eight different chunk columns, the eight writes eight different rows (both conflict free).
    for (int e = threadIdx.x; e < CB * CH; e += blockDim.x) {
        const int blk = e >> 6, j = e & 7, dd = (e >> 3) & 7;
        const int ro = blk / (CH / 8), co = blk % (CH / 8);
        const int r = ro * 8 + j, c = co * 8 + ((dd + j) & 7);
        float4 h, l;
        split_tf32_4(Bp4[r * CH + c], h, l);
        const int o = (r / M) * (M * CH) + c * M + (r % M);
        Ahi[o] = h;
        Alo[o] = l;
    }
    {   // P (symmetric: row n of the B operand is row n of P): chunk (n, c) at (c * KP + n) * 16
        const double* __restrict__ Pinv = D.Pinv;
        for (int e = threadIdx.x; e < KP * CH; e += blockDim.x) {
            const int nn = e % KP, c = e / KP;
            const float4 v = make_float4((float)Pinv[nn * KP + c * 4], (float)Pinv[nn * KP + c * 4 + 1],
                                         (float)Pinv[nn * KP + c * 4 + 2], (float)Pinv[nn * KP + c * 4 + 3]);
            float4 h, l;
            split_tf32_4(v, h, l);
            Phi[c * KP + nn] = h;
            Plo[c * KP + nn] = l;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    if (threadIdx.x == 0) {
        // instruction descriptor: D = F32, A = B = TF32, both K-major, N = KP, M = 128
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        constexpr uint32_t A_LBO = M * 16, P_LBO = KP * 16, SBO = 128, TILE_BYTES = M * KP * 4;
        const uint32_t a_hi = smem_u32(Ahi), a_lo = smem_u32(Alo), p_hi = smem_u32(Phi), p_lo = smem_u32(Plo);
        for (int t = 0; t < ntile; ++t) {
            uint32_t accum = 0;
#pragma unroll
            for (int term = 0; term < 3; ++term) {
                const uint32_t a0 = ((term == 1) ? a_lo : a_hi) + (uint32_t)t * TILE_BYTES;
                const uint32_t p0 = (term == 2) ? p_lo : p_hi;
#pragma unroll
                for (int ks = 0; ks < KP / 8; ++ks) {   // one MMA consumes K = 8 tf32 = two 16-byte chunks
                    umma_tf32(tmem_d + (uint32_t)(t * KP), umma_desc_kmajor(a0 + ks * 2 * A_LBO, A_LBO, SBO),
                              umma_desc_kmajor(p0 + ks * 2 * P_LBO, P_LBO, SBO), idesc, accum);
                    accum = 1;
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"l"(
                         (uint64_t)__cvta_generic_to_shared(bar + 1))
                     : "memory");
    }
    mbar_wait(bar + 1, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {   // 32 warps = 4 lane quadrants x (CB / 128 tiles) x (KP / 16 column slices): each thread reads 16 floats of one row
        const int qd = warp & 3, rest = warp >> 2;
        const int t = rest % ntile, sl = rest / ntile;
        if (sl < KP / 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem_d + ((uint32_t)(qd * 32) << 16) + (uint32_t)(t * KP + sl * 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = t * M + qd * 32 + lane;
            if (row < ncell) {
                float4* dst = reinterpret_cast<float4*>(D.U + (size_t)(c0 + row) * KP + sl * 16);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    dst[jj] = make_float4(__uint_as_float(v[4 * jj]), __uint_as_float(v[4 * jj + 1]), __uint_as_float(v[4 * jj + 2]),
                                          __uint_as_float(v[4 * jj + 3]));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(128) : "memory");
}

void launch_spmm_ltx_tiled(const DsDev* ds, int nds, int total_blocks, int kp, int cb, int gt, bool fuse_u, cudaStream_t st) {
    if (total_blocks <= 0) return;
    const size_t smem = spmm_ltx_smem(kp, cb, gt, fuse_u);
    if (fuse_u) {
        LGR_REQUIRE(cb % 128 == 0 && (cb / 128) * kp == 128, "fused U epilogue: cell block x KP must fill 128 TMEM columns");
        if (kp == 32)
            k_spmm_ltx_tiled<32, true><<<total_blocks, LTX_THREADS, smem, st>>>(ds, nds, cb, gt);
        else
            k_spmm_ltx_tiled<64, true><<<total_blocks, LTX_THREADS, smem, st>>>(ds, nds, cb, gt);
    } else if (kp == 32) {
        k_spmm_ltx_tiled<32, false><<<total_blocks, LTX_THREADS, smem, st>>>(ds, nds, cb, gt);
    } else {
        k_spmm_ltx_tiled<64, false><<<total_blocks, LTX_THREADS, smem, st>>>(ds, nds, cb, gt);
    }
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
    LGR_CUDA(cudaFuncSetAttribute(k_spmm_ltx_tiled<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    LGR_CUDA(cudaFuncSetAttribute(k_spmm_ltx_tiled<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    LGR_CUDA(cudaFuncSetAttribute(k_spmm_ltx_tiled<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    LGR_CUDA(cudaFuncSetAttribute(k_spmm_ltx_tiled<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
}

}  // namespace lgr

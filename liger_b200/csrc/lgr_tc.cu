// liger_b200 -- the dense k x k contraction of the H solve on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   U = B P      B: n x KP rhs rows of the H solve (fp32),  P = CtC^-1 (KP x KP, symmetric),  U: n x KP
//
// u_j = P b_j is what the dual side of the NNLS needs for every column (lgr_nnls.cu); it is the only GEMM-shaped work
// of the iteration (2 k^2 flops per cell).  Single-pass TF32 is not admissible (a 10-bit mantissa breaks
// the 1e-3 / 1e-4 parity gates), so the product is split  x = hi + lo  with hi = x truncated to TF32:
//       U = B_hi P_hi + B_lo P_hi + B_hi P_lo            (3 x TF32, fp32 accumulation in TMEM, error ~2^-21)
//
// One CTA (128 threads) owns tiles of 128 cells:  the B tile is written to shared memory in the canonical K-major
// no-swizzle UMMA layout (16-byte chunk (row r, k-chunk c) at (c * M + r) * 16: SBO = 128 B, LBO = M * 16 B), one
// elected thread issues 12 `tcgen05.mma.cta_group::1.kind::tf32` (M = 128, N = KP, K = 8) into a KP-column TMEM
// accumulator and commits to an mbarrier; the four warps read their 32 TMEM lanes back with `tcgen05.ld.32x32b` and
// store the rows of U.
#include "lgr_common.cuh"
#include "lgr_ptx.cuh"

namespace lgr {

namespace {

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start [0,14) >>4, LBO [16,30) >>4, SBO [32,46) >>4, version [46,48) = 1, layout [61,64) = 0
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);   // keep sign, exponent and the 10 leading mantissa bits
    lo = x - hi;                                              // exact in fp32
}

}  // namespace

template <int KP>
__global__ void __launch_bounds__(128) k_pb_tc(const NnlsGroup* __restrict__ groups, int ngroups, long long total_tiles) {
    constexpr int M = 128;
    constexpr int CH = KP / 4;                        // 16-byte k-chunks per row
    constexpr uint32_t A_BYTES = M * KP * 4;          // one operand tile
    constexpr uint32_t P_BYTES = KP * KP * 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* Ahi = reinterpret_cast<float4*>(smem_raw);
    float4* Alo = reinterpret_cast<float4*>(smem_raw + A_BYTES);
    float4* Phi = reinterpret_cast<float4*>(smem_raw + 2 * A_BYTES);
    float4* Plo = reinterpret_cast<float4*>(smem_raw + 2 * A_BYTES + P_BYTES);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(KP)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) mbar_init(&bar, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = KP, M = 128
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

    uint32_t phase = 0;
    int cur_group = -1;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int g = 0;
        while (g + 1 < ngroups && tile >= groups[g + 1].col_off) ++g;     // col_off counts TILES in this launch
        const NnlsGroup grp = groups[g];
        const int row0 = (int)(tile - grp.col_off) * M;
        const int nrow = min(M, grp.ncols - row0);
        __syncthreads();   // previous tile: all TMEM reads and operand reads are complete
        if (g != cur_group) {
            cur_group = g;
            // P tile (symmetric, so row n of the B operand is row n of P): chunk (n, c) at (c * KP + n) * 16
            for (int e = tid; e < KP * CH; e += 128) {
                const int n = e % KP, c = e / KP;
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) split_tf32((float)grp.ginv[n * KP + c * 4 + j], hi[j], lo[j]);
                Phi[c * KP + n] = make_float4(hi[0], hi[1], hi[2], hi[3]);
                Plo[c * KP + n] = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        // B tile: chunk (r, c) at (c * M + r) * 16; lanes walk r so the shared-memory stores are contiguous
        const float4* __restrict__ B4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(grp.rhs) + (size_t)row0 * KP);
        for (int e = tid; e < M * CH; e += 128) {
            const int r = e % M, c = e / M;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrow) v = __ldg(B4 + (size_t)r * CH + c);
            float4 h, l;
            split_tf32(v.x, h.x, l.x);
            split_tf32(v.y, h.y, l.y);
            split_tf32(v.z, h.z, l.z);
            split_tf32(v.w, h.w, l.w);
            Ahi[c * M + r] = h;
            Alo[c * M + r] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(Ahi), a_lo = smem_u32(Alo), p_hi = smem_u32(Phi), p_lo = smem_u32(Plo);
            // one MMA consumes K = 8 tf32 = two 16-byte chunks: advance the start address by 2 * LBO per step
            constexpr uint32_t A_LBO = M * 16, P_LBO = KP * 16, SBO = 128;
            uint32_t acc = 0;
#pragma unroll
            for (int term = 0; term < 3; ++term) {
                const uint32_t a0 = (term == 1) ? a_lo : a_hi;
                const uint32_t p0 = (term == 2) ? p_lo : p_hi;
#pragma unroll
                for (int ks = 0; ks < KP / 8; ++ks) {
                    tc_mma_tf32(tmem_d, umma_desc(a0 + ks * 2 * A_LBO, A_LBO, SBO), umma_desc(p0 + ks * 2 * P_LBO, P_LBO, SBO),
                                idesc, acc);
                    acc = 1;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"l"(
                             (uint64_t)__cvta_generic_to_shared(&bar))
                         : "memory");
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // epilogue: warp w owns TMEM lanes 32w..32w+31 = rows of the tile; each thread reads its whole row
        float* __restrict__ U = reinterpret_cast<float*>(grp.x) + (size_t)row0 * KP;
        const int r = tid;
#pragma unroll
        for (int c0 = 0; c0 < KP; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
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
            if (r < nrow) {
                float4* dst = reinterpret_cast<float4*>(U + (size_t)r * KP + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                         __uint_as_float(v[4 * j + 3]));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(KP) : "memory");
}

size_t pb_tc_smem(int kp) { return (size_t)2 * 128 * kp * 4 + (size_t)2 * kp * kp * 4; }

// groups[g]: ginv = P, rhs = B (n x KP fp32), x = U (n x KP fp32), ncols = n, col_off = first 128-row tile of the group
void launch_pb_tc(const NnlsGroup* groups, int ngroups, long long total_tiles, int kp, int num_sms, cudaStream_t st) {
    if (total_tiles <= 0) return;
    // several CTAs per SM (40 KB of shared memory and KP TMEM columns each): their load / MMA / epilogue phases overlap
    const long long cap = (long long)num_sms * (kp == 32 ? 5 : 2);
    const int grid = (int)(total_tiles < cap ? total_tiles : cap);
    if (kp == 32)
        k_pb_tc<32><<<grid, 128, pb_tc_smem(kp), st>>>(groups, ngroups, total_tiles);
    else
        k_pb_tc<64><<<grid, 128, pb_tc_smem(kp), st>>>(groups, ngroups, total_tiles);
    LGR_CUDA(cudaGetLastError());
}

void tc_init() {
    LGR_CUDA(cudaFuncSetAttribute(k_pb_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb_tc_smem(32)));
    LGR_CUDA(cudaFuncSetAttribute(k_pb_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb_tc_smem(64)));
}

}  // namespace lgr

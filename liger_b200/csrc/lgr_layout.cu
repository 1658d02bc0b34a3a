// liger_b200 -- one-time layout conversion on the device (upload path, not part of the timed ANLS loop).
//   * fp64 -> fp32 value conversion, ||X||_F^2
//   * stacking [E_i; P_i] for UINMF
//   * the tiled-CSR copy of X used by k_spmm_xh (stable radix sort on the key  tile * rows + gene)
//   * R (column-major fp64) <-> device (row-padded) factor layouts
#include <cub/cub.cuh>

#include "lgr_common.cuh"

namespace lgr {

__global__ void k_f64_to_f32(const double* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (float)in[i];
}
void dev_f64_to_f32(const double* in, float* out, size_t n, cudaStream_t st) {
    if (!n) return;
    const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    k_f64_to_f32<<<grid, 256, 0, st>>>(in, out, n);
    LGR_CUDA(cudaGetLastError());
}

__global__ void k_sumsq(const float* __restrict__ v, size_t n, double* out) {
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double x = (double)v[i];
        acc += x * x;
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out, sh[0]);
}
void dev_sumsq(const float* v, size_t n, double* out, cudaStream_t st) {
    if (!n) return;
    const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    k_sumsq<<<grid, 256, 0, st>>>(v, n, out);
    LGR_CUDA(cudaGetLastError());
}

__global__ void k_add_int(int* __restrict__ p, size_t n, int delta) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] += delta;
}
void dev_add_int(int* p, size_t n, int delta, cudaStream_t st) {
    if (!n) return;
    k_add_int<<<(int)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(p, n, delta);
    LGR_CUDA(cudaGetLastError());
}

// stacked column j = E column j followed by P column j with row indices shifted by m
__global__ void k_stack_ptr(int n, const int* pE, const int* pP, int* pS) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j <= n) pS[j] = pE[j] + pP[j];
}
__global__ void k_stack_fill(int n, const int* pE, const int* iE, const float* vE, const int* pP, const int* iP,
                             const float* vP, int m, const int* pS, int* iS, float* vS) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const int be = pE[warp], ne = pE[warp + 1] - be, bp = pP[warp], np = pP[warp + 1] - bp, o = pS[warp];
    for (int t = lane; t < ne; t += 32) {
        iS[o + t] = iE[be + t];
        vS[o + t] = vE[be + t];
    }
    for (int t = lane; t < np; t += 32) {
        iS[o + ne + t] = iP[bp + t] + m;
        vS[o + ne + t] = vP[bp + t];
    }
}
void dev_stack_csc(int n, const int* pE, const int* iE, const float* vE, const int* pP, const int* iP, const float* vP,
                   int m, int* pS, int* iS, float* vS, cudaStream_t st) {
    k_stack_ptr<<<(n + 1 + 255) / 256, 256, 0, st>>>(n, pE, pP, pS);
    LGR_CUDA(cudaGetLastError());
    if (n > 0) {
        k_stack_fill<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, st>>>(n, pE, iE, vE, pP, iP, vP, m, pS, iS, vS);
        LGR_CUDA(cudaGetLastError());
    }
}

// ---- padded tile formats (see lgr_spmm.cu) ---------------------------------------------------------
//   mode 0, gene-tiled CSC: key = (row / tile) * n + col ,  loc = row % tile      (k_spmm_ltx_tiled)
//   mode 1, cell-tiled CSR: key = (col / tile) * rows + row, loc = col % tile      (k_spmm_xh)
// Entries are stably sorted by key (so the minor index stays ascending inside a segment) and every segment is
// padded with (loc 0, value 0) entries to a multiple of 8; ptr8[key] is the segment start in groups of 8.
// The stream is SELF-DESCRIBING: besides the 16-bit local indices (pre-scaled by KP/4) and the fp32 values, every group
// of 8 has a 32-bit tag = its output row (mode 0: cell % tagmod, mode 1: gene), so that a consumer can be handed ANY
// contiguous range of groups without looking at ptr8.
__global__ void k_tile_keys(int mode, int n, int rows, int tile, int scale, const int* __restrict__ colptr,
                            const int* __restrict__ rowidx, unsigned int* __restrict__ keys,
                            unsigned int* __restrict__ perm, unsigned short* __restrict__ loc, int* __restrict__ counts) {
    const int warp = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= n) return;
    const int b = colptr[warp], e = colptr[warp + 1];
    for (int t = b + lane; t < e; t += 32) {
        const int r = rowidx[t];
        unsigned int key;
        unsigned short l;
        if (mode == 0) {
            key = (unsigned)(r / tile) * (unsigned)n + (unsigned)warp;
            l = (unsigned short)((r % tile) * scale);
        } else {
            key = (unsigned)(warp / tile) * (unsigned)rows + (unsigned)r;
            l = (unsigned short)((warp % tile) * scale);
        }
        keys[t] = key;
        perm[t] = (unsigned)t;
        loc[t] = l;
        atomicAdd(&counts[key], 1);
    }
}
__global__ void k_pad_counts(size_t nkeys, const int* __restrict__ counts, int* __restrict__ groups) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i <= nkeys; i += (size_t)gridDim.x * blockDim.x)
        groups[i] = i < nkeys ? (counts[i] + 7) >> 3 : 0;
}
__global__ void k_tile_scatter(size_t nnz, const unsigned int* __restrict__ keys_sorted, const unsigned int* __restrict__ perm,
                               const int* __restrict__ segstart, const int* __restrict__ ptr8,
                               const unsigned short* __restrict__ loc, const float* __restrict__ val,
                               unsigned int minor, unsigned int tagmod, unsigned short* __restrict__ oloc,
                               unsigned int* __restrict__ otag, float* __restrict__ oval) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned int key = keys_sorted[i], p = perm[i];
        const size_t off = i - (size_t)segstart[key];
        const size_t base = (size_t)ptr8[key] * 8;
        oloc[base + off] = loc[p];
        oval[base + off] = val[p];
        if ((off & 7) == 0) otag[(size_t)ptr8[key] + (off >> 3)] = (key % minor) % tagmod;
    }
}

void dev_build_padded(int mode, int n, int rows, int tile, int scale, int tagmod, const int* colptr, const int* rowidx,
                      const float* val, size_t nnz, DBuf<int>& ptr8, DBuf<unsigned short>& oloc, DBuf<unsigned int>& otag,
                      DBuf<float>& oval, cudaStream_t st) {
    // local indices are stored pre-scaled by `scale` (= KP/4, the float4 slots per staged row)
    LGR_REQUIRE(tile >= 1 && (long long)tile * scale <= 65536, "tile size x KP/4 must fit a 16-bit local index");
    if (mode == 1 || tagmod < 1) tagmod = 0x7fffffff;
    const size_t nmajor = mode == 0 ? (size_t)(rows + tile - 1) / tile : (size_t)(n + tile - 1) / tile;
    const size_t nkeys = nmajor * (size_t)(mode == 0 ? n : rows);
    LGR_REQUIRE(nkeys < (size_t)0x7fffffff, "tiled format: tiles x minor dimension exceeds 2^31");
    ptr8.alloc(nkeys + 1);
    ptr8.zero(st);
    if (nnz == 0 || n == 0) {
        oloc.alloc(8);
        oval.alloc(8);
        otag.alloc(1);
        oloc.zero(st);
        oval.zero(st);
        otag.zero(st);
        return;
    }
    DBuf<unsigned int> keys(nnz), keys2(nnz), perm(nnz), perm2(nnz);
    DBuf<unsigned short> loc(nnz);
    DBuf<int> counts(nkeys + 1), segstart(nkeys + 1), groups(nkeys + 1);
    counts.zero(st);
    k_tile_keys<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, st>>>(mode, n, rows, tile, scale, colptr, rowidx, keys.p, perm.p,
                                                                             loc.p, counts.p);
    LGR_CUDA(cudaGetLastError());
    int bits = 1;
    while (((size_t)1 << bits) < nkeys) ++bits;
    size_t tmp_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys2.p, perm.p, perm2.p, (int)nnz, 0, bits, st);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, counts.p, segstart.p, (int)(nkeys + 1), st);
    if (scan_bytes > tmp_bytes) tmp_bytes = scan_bytes;
    DBuf<unsigned char> tmp(tmp_bytes);
    LGR_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys2.p, perm.p, perm2.p, (int)nnz, 0, bits, st));
    LGR_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, counts.p, segstart.p, (int)(nkeys + 1), st));
    const int grid = (int)std::min<size_t>((nkeys + 256) / 256, 148 * 16);
    k_pad_counts<<<grid, 256, 0, st>>>(nkeys, counts.p, groups.p);
    LGR_CUDA(cudaGetLastError());
    LGR_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, groups.p, ptr8.p, (int)(nkeys + 1), st));
    int total_groups = 0;
    LGR_CUDA(cudaMemcpyAsync(&total_groups, ptr8.p + nkeys, sizeof(int), cudaMemcpyDeviceToHost, st));
    LGR_CUDA(cudaStreamSynchronize(st));
    const size_t padded = (size_t)total_groups * 8 + 8;
    oloc.alloc(padded);
    oval.alloc(padded);
    otag.alloc((size_t)total_groups + 1);
    oloc.zero(st);
    oval.zero(st);
    otag.zero(st);
    const int grid2 = (int)std::min<size_t>((nnz + 255) / 256, 148 * 16);
    k_tile_scatter<<<grid2, 256, 0, st>>>(nnz, keys2.p, perm2.p, segstart.p, ptr8.p, loc.p, val, (unsigned)(mode == 0 ? n : rows),
                                         (unsigned)tagmod, oloc.p, otag.p, oval.p);
    LGR_CUDA(cudaGetLastError());
    LGR_CUDA(cudaStreamSynchronize(st));
}

void dev_exclusive_scan_int(const int* in, int* out, size_t n, cudaStream_t st) {
    if (!n) return;
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, st);
    DBuf<unsigned char> tmp(bytes);
    LGR_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)n, st));
}

// ---- factor layout conversion -------------------------------------------------------------------
__global__ void k_colmajor_to_padded(const double* __restrict__ in, double* __restrict__ out, int rows, int k, int kp) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * kp) return;
    const int r = (int)(idx / kp), f = (int)(idx % kp);
    out[idx] = f < k ? in[(size_t)f * rows + r] : 0.0;
}
__global__ void k_padded_to_colmajor(const double* __restrict__ in, double* __restrict__ out, int rows, int k, int kp) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * k) return;
    const int f = (int)(idx / rows), r = (int)(idx % rows);
    out[idx] = in[(size_t)r * kp + f];
}
__global__ void k_hpad_to_colmajor(const float* __restrict__ H, double* __restrict__ out, int n, int k, int kp) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * k) return;
    const int f = (int)(idx / n), c = (int)(idx % n);
    out[idx] = (double)H[(size_t)c * kp + f];
}
__global__ void k_colmajor_to_hpad(const double* __restrict__ in, float* __restrict__ H, int n, int k, int kp) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * kp) return;
    const int c = (int)(idx / kp), f = (int)(idx % kp);
    H[idx] = f < k ? (float)in[(size_t)f * n + c] : 0.f;
}
__global__ void k_mask_from_positive(const float* __restrict__ H, unsigned long long* __restrict__ mask, int n, int k, int kp) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    unsigned long long m = 0;
    for (int f = 0; f < k; ++f)
        if (H[(size_t)c * kp + f] > 0.f) m |= 1ull << f;
    mask[c] = m;
}
#define LGR_GRID(total) (unsigned)(((size_t)(total) + 255) / 256)
void dev_colmajor_to_padded(const double* in, double* out, int rows, int k, int kp, cudaStream_t st) {
    if (!rows) return;
    k_colmajor_to_padded<<<LGR_GRID((size_t)rows * kp), 256, 0, st>>>(in, out, rows, k, kp);
    LGR_CUDA(cudaGetLastError());
}
void dev_padded_to_colmajor(const double* in, double* out, int rows, int k, int kp, cudaStream_t st) {
    if (!rows) return;
    k_padded_to_colmajor<<<LGR_GRID((size_t)rows * k), 256, 0, st>>>(in, out, rows, k, kp);
    LGR_CUDA(cudaGetLastError());
}
void dev_hpad_to_colmajor(const float* H, double* out, int n, int k, int kp, cudaStream_t st) {
    if (!n) return;
    k_hpad_to_colmajor<<<LGR_GRID((size_t)n * k), 256, 0, st>>>(H, out, n, k, kp);
    LGR_CUDA(cudaGetLastError());
}
void dev_colmajor_to_hpad(const double* in, float* H, int n, int k, int kp, cudaStream_t st) {
    if (!n) return;
    k_colmajor_to_hpad<<<LGR_GRID((size_t)n * kp), 256, 0, st>>>(in, H, n, k, kp);
    LGR_CUDA(cudaGetLastError());
}
void dev_mask_from_positive(const float* H, unsigned long long* mask, int n, int k, int kp, cudaStream_t st) {
    if (!n) return;
    k_mask_from_positive<<<LGR_GRID(n), 256, 0, st>>>(H, mask, n, k, kp);
    LGR_CUDA(cudaGetLastError());
}

}  // namespace lgr

// liger_b200 -- the native pieces of the preprocessing that produces the path's input (scaleData), on the device.
// "the step immediately before the path".  Pure streaming scans over a CSC matrix (HBM bound):
//
//   k_col_normalize      x / colSums(x)[col]                 normalize.dgCMatrix        (R/preprocess.R:592-608)
//   k_row_sum_count      rowSums, non-zeros per row          Matrix::rowMeans           (R/preprocess.R:1024)
//   k_row_sqdev          sum_nz (x - mean_row)^2             rowVars_sparse_rcpp        (src/data_processing.cpp:152-172)
//   k_row_sumsq          sum_nz x^2                          scaleNotCenter_byRow_rcpp  (src/data_processing.cpp:42-58)
//   k_row_scale          x / sqrt(sumsq_row / (ncol - 1)), 0 where that is 0
//
// All arithmetic is FP64 like the reference.  Row reductions over a cell-major (CSC) matrix are scatters: they use
// fire-and-forget fp64 `red.global.add` (one per non-zero; the row vectors live in L2).  Summation order differs from
// the reference's serial loops, so results agree to ~1e-15 relative, not bitwise.
#include "lgr_common.cuh"

namespace lgr {

// one warp per column: sum in fp64 (lane-strided, then a shuffle tree), then scale the column
__global__ void __launch_bounds__(256) k_col_normalize(int ncol, const int* __restrict__ colptr, const double* __restrict__ x,
                                                       double* __restrict__ out, double* __restrict__ colsum) {
    const int warp = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= ncol) return;
    const int b = colptr[warp], e = colptr[warp + 1];
    double s = 0.0;
    for (int t = b + lane; t < e; t += 32) s += x[t];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0 && colsum) colsum[warp] = s;
    for (int t = b + lane; t < e; t += 32) out[t] = x[t] / s;
}

__global__ void __launch_bounds__(256) k_row_sum_count(size_t nnz, const int* __restrict__ rowidx, const double* __restrict__ x,
                                                       double* __restrict__ rsum, double* __restrict__ rcount) {
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (size_t)gridDim.x * blockDim.x) {
        const int r = rowidx[t];
        atomicAdd(rsum + r, x[t]);
        atomicAdd(rcount + r, 1.0);
    }
}

__global__ void __launch_bounds__(256) k_row_sqdev(size_t nnz, const int* __restrict__ rowidx, const double* __restrict__ x,
                                                   const double* __restrict__ means, double* __restrict__ acc) {
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (size_t)gridDim.x * blockDim.x) {
        const int r = rowidx[t];
        const double dlt = x[t] - means[r];
        atomicAdd(acc + r, dlt * dlt);
    }
}

// vars = (sqdev + (ncol_x - count) * mean^2) / (ncol - 1)      (src/data_processing.cpp:166-170)
__global__ void k_row_var_finish(int nrow, double ncol_x, double ncol, const double* __restrict__ means,
                                 const double* __restrict__ count, double* __restrict__ vars) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrow) return;
    const double m2 = means[r] * means[r];
    vars[r] = (vars[r] + (ncol_x - count[r]) * m2) / (ncol - 1.0);
}

__global__ void k_row_mean_finish(int nrow, double ncol_x, const double* __restrict__ rsum, double* __restrict__ means) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nrow) means[r] = rsum[r] / ncol_x;
}

__global__ void __launch_bounds__(256) k_row_sumsq(size_t nnz, const int* __restrict__ rowidx, const double* __restrict__ x,
                                                   double* __restrict__ ss) {
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (size_t)gridDim.x * blockDim.x) {
        const double v = x[t];
        atomicAdd(ss + rowidx[t], v * v);
    }
}

__global__ void k_row_rms_finish(int nrow, double ncol, double* __restrict__ ss) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nrow) ss[r] = sqrt(ss[r] / (ncol - 1.0));
}

__global__ void __launch_bounds__(256) k_row_scale(size_t nnz, const int* __restrict__ rowidx, const double* __restrict__ x,
                                                   const double* __restrict__ rms, double* __restrict__ out) {
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (size_t)gridDim.x * blockDim.x) {
        const double s = rms[rowidx[t]];
        out[t] = (s == 0.0) ? 0.0 : x[t] / s;
    }
}

// ---- row subset of a CSC matrix (scaleNotCenter.dgCMatrix: object[features, ]) -------------------------------------
// lut[row] = new row index (position in the feature list) or -1.  One warp per column.
__global__ void __launch_bounds__(256) k_subset_count(int ncol, const int* __restrict__ colptr, const int* __restrict__ rowidx,
                                                      const int* __restrict__ lut, int* __restrict__ count) {
    const int warp = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= ncol) return;
    int c = 0;
    for (int t = colptr[warp] + lane; t < colptr[warp + 1]; t += 32) c += lut[rowidx[t]] >= 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if (lane == 0) count[warp] = c;
}
__global__ void __launch_bounds__(256) k_subset_fill(int ncol, const int* __restrict__ colptr, const int* __restrict__ rowidx,
                                                     const double* __restrict__ x, const int* __restrict__ lut,
                                                     const int* __restrict__ newptr, int* __restrict__ newrow,
                                                     double* __restrict__ newx) {
    const int warp = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= ncol) return;
    int o = newptr[warp];
    const int b = colptr[warp], e = colptr[warp + 1];
    for (int t0 = b; t0 < e; t0 += 32) {   // warp-uniform trip count; order inside the column is preserved
        const int t = t0 + lane;
        const int nr = (t < e) ? lut[rowidx[t]] : -1;
        const unsigned keep = __ballot_sync(0xffffffffu, nr >= 0);
        if (nr >= 0) {
            const int dst = o + __popc(keep & ((1u << lane) - 1u));
            newrow[dst] = nr;
            newx[dst] = x[t];
        }
        o += __popc(keep);
    }
}
__global__ void k_f64_to_f32_prep(const double* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (float)in[i];
}

__global__ void k_check_rows(size_t nnz, const int* __restrict__ rowidx, int nrow, int* __restrict__ bad) {
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (size_t)gridDim.x * blockDim.x)
        if ((unsigned)rowidx[t] >= (unsigned)nrow) *bad = 1;
}

static int stream_grid(size_t n) { return (int)std::min<size_t>((n + 255) / 256, (size_t)148 * 16); }

void prep_col_normalize(int ncol, const int* colptr, const double* x, double* out, double* colsum, cudaStream_t st) {
    if (ncol <= 0) return;
    k_col_normalize<<<(unsigned)(((size_t)ncol * 32 + 255) / 256), 256, 0, st>>>(ncol, colptr, x, out, colsum);
    LGR_CUDA(cudaGetLastError());
}

// means_in == nullptr: means = rowSums / ncol_x are computed here.  means_out / vars_out: nrow doubles on the device.
void prep_row_mean_var(int nrow, int ncol_x, size_t nnz, const int* rowidx, const double* x, const double* means_in,
                       double ncol, double* means_out, double* vars_out, double* scratch_count, cudaStream_t st) {
    if (nrow <= 0) return;
    LGR_CUDA(cudaMemsetAsync(vars_out, 0, (size_t)nrow * sizeof(double), st));
    LGR_CUDA(cudaMemsetAsync(scratch_count, 0, (size_t)nrow * sizeof(double), st));
    LGR_CUDA(cudaMemsetAsync(means_out, 0, (size_t)nrow * sizeof(double), st));
    const int rb = (nrow + 255) / 256;
    if (nnz) {
        k_row_sum_count<<<stream_grid(nnz), 256, 0, st>>>(nnz, rowidx, x, means_out, scratch_count);
        LGR_CUDA(cudaGetLastError());
    }
    if (means_in)
        LGR_CUDA(cudaMemcpyAsync(means_out, means_in, (size_t)nrow * sizeof(double), cudaMemcpyDeviceToDevice, st));
    else
        k_row_mean_finish<<<rb, 256, 0, st>>>(nrow, (double)ncol_x, means_out, means_out);
    if (nnz) {
        k_row_sqdev<<<stream_grid(nnz), 256, 0, st>>>(nnz, rowidx, x, means_out, vars_out);
        LGR_CUDA(cudaGetLastError());
    }
    k_row_var_finish<<<rb, 256, 0, st>>>(nrow, (double)ncol_x, ncol, means_out, scratch_count, vars_out);
    LGR_CUDA(cudaGetLastError());
}

void prep_scale_rows(int nrow, int ncol, size_t nnz, const int* rowidx, const double* x, double* out, double* rms,
                     cudaStream_t st) {
    if (nrow <= 0) return;
    LGR_CUDA(cudaMemsetAsync(rms, 0, (size_t)nrow * sizeof(double), st));
    if (nnz) {
        k_row_sumsq<<<stream_grid(nnz), 256, 0, st>>>(nnz, rowidx, x, rms);
        LGR_CUDA(cudaGetLastError());
    }
    k_row_rms_finish<<<(nrow + 255) / 256, 256, 0, st>>>(nrow, (double)ncol, rms);
    LGR_CUDA(cudaGetLastError());
    if (nnz) {
        k_row_scale<<<stream_grid(nnz), 256, 0, st>>>(nnz, rowidx, x, rms, out);
        LGR_CUDA(cudaGetLastError());
    }
}


void prep_subset_count(int ncol, const int* colptr, const int* rowidx, const int* lut, int* count, cudaStream_t st) {
    if (ncol <= 0) return;
    k_subset_count<<<(unsigned)(((size_t)ncol * 32 + 255) / 256), 256, 0, st>>>(ncol, colptr, rowidx, lut, count);
    LGR_CUDA(cudaGetLastError());
}
void prep_subset_fill(int ncol, const int* colptr, const int* rowidx, const double* x, const int* lut, const int* newptr,
                      int* newrow, double* newx, cudaStream_t st) {
    if (ncol <= 0) return;
    k_subset_fill<<<(unsigned)(((size_t)ncol * 32 + 255) / 256), 256, 0, st>>>(ncol, colptr, rowidx, x, lut, newptr, newrow, newx);
    LGR_CUDA(cudaGetLastError());
}
// raw counts usable by the dense-tile sweeps: every stored value an integer in 1..256 (exact in bf16)
__global__ void __launch_bounds__(256) k_check_counts(size_t nnz, const double* __restrict__ x, int* __restrict__ bad) {
    int b = 0;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (size_t)gridDim.x * blockDim.x) {
        const double v = x[t];
        b |= !(v >= 1.0 && v <= 256.0 && v == rint(v));
    }
    if (b) atomicOr(bad, 1);
}
void prep_check_counts(size_t nnz, const double* x, int* bad, cudaStream_t st) {
    if (!nnz) return;
    k_check_counts<<<(int)std::min<size_t>((nnz + 255) / 256, 148 * 16), 256, 0, st>>>(nnz, x, bad);
    LGR_CUDA(cudaGetLastError());
}
__global__ void k_reciprocal(size_t n, const double* __restrict__ in, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] == 0.0 ? 0.0 : 1.0 / in[i];
}
__global__ void k_check_same(size_t n, const double* __restrict__ a, const double* __restrict__ b, int* __restrict__ bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && a[i] != b[i]) atomicOr(bad, 1);
}
void prep_check_same(size_t n, const double* a, const double* b, int* bad, cudaStream_t st) {
    if (!n) return;
    k_check_same<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, a, b, bad);
    LGR_CUDA(cudaGetLastError());
}
void prep_reciprocal(size_t n, const double* in, double* out, cudaStream_t st) {
    if (!n) return;
    k_reciprocal<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, in, out);
    LGR_CUDA(cudaGetLastError());
}

void prep_check_rows(size_t nnz, const int* rowidx, int nrow, int* bad, cudaStream_t st) {
    if (!nnz) return;
    k_check_rows<<<stream_grid(nnz), 256, 0, st>>>(nnz, rowidx, nrow, bad);
    LGR_CUDA(cudaGetLastError());
}
void prep_f64_to_f32(const double* in, float* out, size_t n, cudaStream_t st) {
    if (!n) return;
    k_f64_to_f32_prep<<<stream_grid(n), 256, 0, st>>>(in, out, n);
    LGR_CUDA(cudaGetLastError());
}

}  // namespace lgr

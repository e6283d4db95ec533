// Batched truncated SVD for LoRA distillation (replaces the serial torch.linalg.svd loop of
// /root/reference/lora_diffusion/cli_svd.py:24-92).
//
// The reference takes a FULL SVD of every weight delta dW = W_tuned - W_base (up to 10240x1280,
// conv deltas up to 1280x23040 with full_matrices=True) and keeps the top r <= 16 triplets. Here the
// top-r subspace is found with a randomized range finder (Halko-Martinsson-Tropp) with L = 32
// probe vectors (oversampling 32 - r >= 16) and q power iterations, every pass re-orthonormalised
// by CholeskyQR2; the small LxL problem is solved by cyclic Jacobi inside one warp. All passes
// stream dW = W_t - W_b straight from the two stored weight tensors (no fp32 delta is ever
// materialised): the kernels are HBM-bound, bytes per pass = 2 * N*K * sizeof(weight).
//
// Primitives (each batched over `batch` same-shape matrices through pointer arrays; tall-skinny
// operands are fp32 row-major [rows, 32]):
//   lb_svd_mul        Y[b] = dW[b] . Z[b]        ([N,K].[K,32])   or   Z[b] = dW[b]^T . Y[b]
//   lb_svd_gram       G[b] = Y[b]^T Y[b]                                  (32x32, atomics)
//   lb_svd_chol_inv   G = R^T R  ->  Rinv                                 (one warp per matrix)
//   lb_svd_apply      Y[b] <- Y[b] . M[b] (* column scale)                ([rows,32].[32,32])
//   lb_svd_jacobi     G = V diag(w) V^T, w sorted descending              (one warp per matrix)
//   lb_svd_randn      probe matrix from a counter hash
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "lora_b200.h"

namespace lbsvd {

constexpr int L = 32;  // probe / subspace width

__device__ __forceinline__ float ld_w(const void* p, long long i, int dt) {
  if (dt == LB_F32) return reinterpret_cast<const float*>(p)[i];
  if (dt == LB_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}

// ---------------------------------------------------------------------------------- Y = dW . Z
// CTA: 64 rows of dW x all 32 columns; K walked in chunks of 32. 256 threads, thread = 4 rows x 2 cols.
__global__ void __launch_bounds__(256)
mul_right_kernel(const void* const* __restrict__ Wt, const void* const* __restrict__ Wb, int wdt,
                 const float* __restrict__ Z, float* __restrict__ Y, int N, int K) {
  __shared__ float As[64][33];
  __shared__ float Zs[32][L];
  const int b = blockIdx.y;
  const void* wt = Wt[b];
  const void* wb = Wb ? Wb[b] : nullptr;
  const float* z = Z + static_cast<size_t>(b) * K * L;
  float* y = Y + static_cast<size_t>(b) * N * L;
  const int n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // cols 2*tx..+1, rows 4*ty..+3
  float acc[4][2] = {};
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int i = threadIdx.x; i < 64 * 32; i += 256) {
      const int rr = i >> 5, cc = i & 31;
      const int n = n0 + rr, k = k0 + cc;
      float v = 0.f;
      if (n < N && k < K) {
        const long long idx = static_cast<long long>(n) * K + k;
        v = ld_w(wt, idx, wdt) - (wb ? ld_w(wb, idx, wdt) : 0.f);
      }
      As[rr][cc] = v;
    }
    for (int i = threadIdx.x; i < 32 * L; i += 256) {
      const int kk = i >> 5, cc = i & 31;
      Zs[kk][cc] = (k0 + kk < K) ? z[static_cast<size_t>(k0 + kk) * L + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const float z0 = Zs[kk][2 * tx], z1 = Zs[kk][2 * tx + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = As[4 * ty + i][kk];
        acc[i][0] += a * z0;
        acc[i][1] += a * z1;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + 4 * ty + i;
    if (n < N) {
      y[static_cast<size_t>(n) * L + 2 * tx] = acc[i][0];
      y[static_cast<size_t>(n) * L + 2 * tx + 1] = acc[i][1];
    }
  }
}

// ---------------------------------------------------------------------------------- Z = dW^T . Y
// CTA: 64 columns (k) of dW x a slab of rows; partial sums are combined with atomics (Z zeroed first).
__global__ void __launch_bounds__(256)
mul_left_kernel(const void* const* __restrict__ Wt, const void* const* __restrict__ Wb, int wdt,
                const float* __restrict__ Y, float* __restrict__ Z, int N, int K, int rows_per_cta) {
  __shared__ float As[32][65];
  __shared__ float Ys[32][L];
  const int b = blockIdx.z;
  const void* wt = Wt[b];
  const void* wb = Wb ? Wb[b] : nullptr;
  const float* y = Y + static_cast<size_t>(b) * N * L;
  float* z = Z + static_cast<size_t>(b) * K * L;
  const int k0 = blockIdx.x * 64;
  const int n_begin = blockIdx.y * rows_per_cta;
  const int n_end = min(N, n_begin + rows_per_cta);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // cols 2*tx..+1, k rows 4*ty..+3
  float acc[4][2] = {};
  for (int nb = n_begin; nb < n_end; nb += 32) {
    for (int i = threadIdx.x; i < 32 * 64; i += 256) {
      const int rr = i >> 6, cc = i & 63;
      const int n = nb + rr, k = k0 + cc;
      float v = 0.f;
      if (n < n_end && k < K) {
        const long long idx = static_cast<long long>(n) * K + k;
        v = ld_w(wt, idx, wdt) - (wb ? ld_w(wb, idx, wdt) : 0.f);
      }
      As[rr][cc] = v;
    }
    for (int i = threadIdx.x; i < 32 * L; i += 256) {
      const int rr = i >> 5, cc = i & 31;
      Ys[rr][cc] = (nb + rr < n_end) ? y[static_cast<size_t>(nb + rr) * L + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      const float y0 = Ys[rr][2 * tx], y1 = Ys[rr][2 * tx + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = As[rr][4 * ty + i];
        acc[i][0] += a * y0;
        acc[i][1] += a * y1;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + 4 * ty + i;
    if (k < K) {
      atomicAdd(z + static_cast<size_t>(k) * L + 2 * tx, acc[i][0]);
      atomicAdd(z + static_cast<size_t>(k) * L + 2 * tx + 1, acc[i][1]);
    }
  }
}

// ---------------------------------------------------------------------------------- G = Y^T Y
__global__ void __launch_bounds__(256)
gram_kernel(const float* __restrict__ Y, float* __restrict__ G, int rows, int rows_per_cta) {
  __shared__ float Ys[64][L + 1];
  const int b = blockIdx.y;
  const float* y = Y + static_cast<size_t>(b) * rows * L;
  float* g = G + static_cast<size_t>(b) * L * L;
  const int r_begin = blockIdx.x * rows_per_cta, r_end = min(rows, r_begin + rows_per_cta);
  // thread -> 4 entries of G: (i, j0..j0+3)
  const int i = threadIdx.x >> 3, j0 = (threadIdx.x & 7) * 4;
  float acc[4] = {};
  for (int rb = r_begin; rb < r_end; rb += 64) {
    for (int t = threadIdx.x; t < 64 * L; t += 256) {
      const int rr = t >> 5, cc = t & 31;
      Ys[rr][cc] = (rb + rr < r_end) ? y[static_cast<size_t>(rb + rr) * L + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 64; ++rr) {
      const float a = Ys[rr][i];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] += a * Ys[rr][j0 + q];
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) atomicAdd(g + i * L + j0 + q, acc[q]);
}

// ---------------------------------------------------------------------------------- Cholesky + inverse
// One warp per matrix: G = R^T R (upper R), Rinv = R^{-1}. A tiny ridge keeps G positive definite
// when a probe direction is numerically dependent.
__global__ void __launch_bounds__(32)
chol_inv_kernel(const float* __restrict__ G, float* __restrict__ Rinv) {
  __shared__ float A[L][L + 1];
  __shared__ float Ri[L][L + 1];
  const int b = blockIdx.x, lane = threadIdx.x;
  const float* g = G + static_cast<size_t>(b) * L * L;
  float tr = 0.f;
  for (int i = 0; i < L; ++i) {
    A[i][lane] = g[i * L + lane];
    if (i == lane) tr = A[i][lane];
  }
  for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
  __syncwarp();
  const float ridge = 1e-10f * tr + 1e-30f;
  // right-looking Cholesky, lane = column
  for (int k = 0; k < L; ++k) {
    const float d = sqrtf(fmaxf(A[k][k] + ridge, 1e-30f));
    __syncwarp();
    if (lane >= k) A[k][lane] = (lane == k) ? d : A[k][lane] / d;   // row k of R
    __syncwarp();
    for (int i = k + 1; i < L; ++i)
      if (lane >= i) A[i][lane] -= A[k][i] * A[k][lane];
    __syncwarp();
  }
  // Rinv by back substitution: column `lane` of Rinv solves R x = e_lane
  for (int i = 0; i < L; ++i) Ri[i][lane] = 0.f;
  __syncwarp();
  for (int i = L - 1; i >= 0; --i) {
    float s = (i == lane) ? 1.f : 0.f;
    for (int j = i + 1; j < L; ++j) s -= A[i][j] * Ri[j][lane];
    Ri[i][lane] = (i <= lane) ? s / A[i][i] : 0.f;
    __syncwarp();
  }
  float* out = Rinv + static_cast<size_t>(b) * L * L;
  for (int i = 0; i < L; ++i) out[i * L + lane] = Ri[i][lane];
}

// ---------------------------------------------------------------------------------- Y <- Y . M (* scale_j)
// out may alias Y (each thread reads its whole row before writing). out_cols <= 32 columns are
// written with row pitch out_pitch; optional transposed write (out[j*out_pitch + row]).
__global__ void __launch_bounds__(256)
apply_kernel(const float* __restrict__ Y, const float* __restrict__ Mx, const float* __restrict__ colscale,
             int scale_mode, float* __restrict__ out, int rows, int out_cols, long long out_pitch,
             int transposed, long long out_batch_stride) {
  __shared__ float Ms[L][L];
  const int b = blockIdx.y;
  const float* y = Y + static_cast<size_t>(b) * rows * L;
  const float* m = Mx + static_cast<size_t>(b) * L * L;
  for (int t = threadIdx.x; t < L * L; t += 256) Ms[t >> 5][t & 31] = m[t];
  __syncthreads();
  const int row = blockIdx.x * 256 + threadIdx.x;
  if (row >= rows) return;
  float v[L];
#pragma unroll
  for (int i = 0; i < L; ++i) v[i] = y[static_cast<size_t>(row) * L + i];
  float* o = out + static_cast<size_t>(b) * out_batch_stride;
  for (int j = 0; j < out_cols; ++j) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < L; ++i) s += v[i] * Ms[i][j];
    if (scale_mode != 0) {
      // colscale is sorted descending (lb_svd_jacobi): entry 0 is the largest singular value.
      // Directions below 1e-6 of it are numerically null: dropped (zero column) instead of divided.
      const float c = colscale[static_cast<size_t>(b) * L + j];
      const float cmax = colscale[static_cast<size_t>(b) * L];
      s = (scale_mode == 1) ? s * c : ((c > 1e-6f * cmax && c > 0.f) ? s / c : 0.f);
    }
    if (transposed) o[static_cast<size_t>(j) * out_pitch + row] = s;
    else o[static_cast<size_t>(row) * out_pitch + j] = s;
  }
}

// ---------------------------------------------------------------------------------- Jacobi eigen
// One warp per symmetric 32x32 matrix: cyclic Jacobi with the round-robin (chess tournament)
// ordering, 16 disjoint rotations per step handled by 16 lane pairs. Outputs eigenvectors
// (columns, sorted by descending eigenvalue) and sqrt(max(eig,0)) as singular values.
__global__ void __launch_bounds__(32)
jacobi_kernel(const float* __restrict__ G, float* __restrict__ V, float* __restrict__ sigma, int sweeps) {
  __shared__ float A[L][L + 1];
  __shared__ float Q[L][L + 1];
  __shared__ int perm[L];
  __shared__ float cs[16], sn[16];
  __shared__ int pp[16], qq[16];
  const int b = blockIdx.x, lane = threadIdx.x;
  const float* g = G + static_cast<size_t>(b) * L * L;
  for (int i = 0; i < L; ++i) {
    A[i][lane] = g[i * L + lane];
    Q[i][lane] = (i == lane) ? 1.f : 0.f;
  }
  perm[lane] = lane;
  __syncwarp();
  for (int sw = 0; sw < sweeps; ++sw) {
    for (int step = 0; step < L - 1; ++step) {
      if (lane < 16) {
        // tournament pairing: positions (i, 31-i) of the current permutation
        int p = perm[lane], q = perm[L - 1 - lane];
        if (p > q) { const int t = p; p = q; q = t; }
        const float apq = A[p][q], app = A[p][p], aqq = A[q][q];
        float c = 1.f, s = 0.f;
        if (fabsf(apq) > 1e-30f) {
          const float tau = (aqq - app) / (2.f * apq);
          const float t = (tau >= 0.f ? 1.f : -1.f) / (fabsf(tau) + sqrtf(1.f + tau * tau));
          c = rsqrtf(1.f + t * t);
          s = t * c;
        }
        cs[lane] = c; sn[lane] = s; pp[lane] = p; qq[lane] = q;
      }
      __syncwarp();
      // A <- J^T A J : first columns (lane = row index), then rows (lane = column index)
      for (int k = 0; k < 16; ++k) {
        const int p = pp[k], q = qq[k];
        const float c = cs[k], s = sn[k];
        const float aip = A[lane][p], aiq = A[lane][q];
        A[lane][p] = c * aip - s * aiq;
        A[lane][q] = s * aip + c * aiq;
        const float vip = Q[lane][p], viq = Q[lane][q];
        Q[lane][p] = c * vip - s * viq;
        Q[lane][q] = s * vip + c * viq;
      }
      __syncwarp();
      for (int k = 0; k < 16; ++k) {
        const int p = pp[k], q = qq[k];
        const float c = cs[k], s = sn[k];
        const float apj = A[p][lane], aqj = A[q][lane];
        A[p][lane] = c * apj - s * aqj;
        A[q][lane] = s * apj + c * aqj;
      }
      __syncwarp();
      // rotate the tournament: position 0 fixed, others shift by one
      int nxt = perm[lane];
      if (lane >= 1) nxt = perm[lane == 1 ? L - 1 : lane - 1];
      __syncwarp();
      perm[lane] = nxt;
      __syncwarp();
    }
  }
  // sort eigenvalues descending (rank by counting), write sorted eigenvectors
  const float w = A[lane][lane];
  int rank = 0;
  for (int j = 0; j < L; ++j) {
    const float wj = A[j][j];
    rank += (wj > w) || (wj == w && j < lane);
  }
  sigma[static_cast<size_t>(b) * L + rank] = sqrtf(fmaxf(w, 0.f));
  float* v = V + static_cast<size_t>(b) * L * L;
  for (int i = 0; i < L; ++i) v[i * L + rank] = Q[i][lane];
}

// ---------------------------------------------------------------------------------- probes
__global__ void randn_kernel(float* __restrict__ out, long long n, unsigned long long seed) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  auto mix = [](unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  };
  const unsigned long long a = mix(seed + 2ull * i * 0x9E3779B97F4A7C15ull);
  const unsigned long long c = mix(seed + (2ull * i + 1ull) * 0x9E3779B97F4A7C15ull);
  const float u1 = (static_cast<float>(a >> 40) + 1.f) * (1.f / 16777217.f);
  const float u2 = static_cast<float>(c >> 40) * (1.f / 16777216.f);
  out[i] = sqrtf(-2.f * __logf(u1)) * __cosf(6.2831853f * u2);
}

}  // namespace lbsvd

using namespace lbsvd;

static inline bool wdt_ok(int dt) { return dt == LB_F32 || dt == LB_BF16 || dt == LB_F16; }

extern "C" int lb_svd_mul(const void* const* Wt, const void* const* Wb, int w_dtype, const float* in,
                          float* out, int N, int K, int batch, int transpose, void* stream) {
  if (N <= 0 || K <= 0 || batch <= 0) return LB_ERR_SHAPE;
  if (!wdt_ok(w_dtype)) return LB_ERR_DTYPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!transpose) {
    dim3 grid((N + 63) / 64, batch);
    mul_right_kernel<<<grid, 256, 0, st>>>(Wt, Wb, w_dtype, in, out, N, K);
  } else {
    if (cudaMemsetAsync(out, 0, sizeof(float) * static_cast<size_t>(batch) * K * L, st) != cudaSuccess)
      return LB_ERR_CUDA;
    int slabs = (N + 1023) / 1024;   // ~1024 rows per CTA
    if (slabs < 1) slabs = 1;
    const int rows_per = ((N + slabs - 1) / slabs + 31) / 32 * 32;
    dim3 grid((K + 63) / 64, (N + rows_per - 1) / rows_per, batch);
    mul_left_kernel<<<grid, 256, 0, st>>>(Wt, Wb, w_dtype, in, out, N, K, rows_per);
  }
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

extern "C" int lb_svd_gram(const float* Y, float* G, int rows, int batch, void* stream) {
  if (rows <= 0 || batch <= 0) return LB_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(G, 0, sizeof(float) * static_cast<size_t>(batch) * L * L, st) != cudaSuccess)
    return LB_ERR_CUDA;
  const int rows_per = 512;
  dim3 grid((rows + rows_per - 1) / rows_per, batch);
  gram_kernel<<<grid, 256, 0, st>>>(Y, G, rows, rows_per);
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

extern "C" int lb_svd_chol_inv(const float* G, float* Rinv, int batch, void* stream) {
  if (batch <= 0) return LB_ERR_SHAPE;
  chol_inv_kernel<<<batch, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(G, Rinv);
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

extern "C" int lb_svd_apply(const float* Y, const float* Mx, const float* colscale, int scale_mode,
                            float* out, int rows, int out_cols, long long out_pitch, int transposed,
                            long long out_batch_stride, int batch, void* stream) {
  if (rows <= 0 || batch <= 0 || out_cols < 1 || out_cols > L) return LB_ERR_SHAPE;
  if (scale_mode != 0 && colscale == nullptr) return LB_ERR_SHAPE;
  dim3 grid((rows + 255) / 256, batch);
  apply_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      Y, Mx, colscale, scale_mode, out, rows, out_cols, out_pitch, transposed, out_batch_stride);
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

extern "C" int lb_svd_jacobi(const float* G, float* V, float* sigma, int batch, int sweeps, void* stream) {
  if (batch <= 0 || sweeps < 1) return LB_ERR_SHAPE;
  jacobi_kernel<<<batch, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(G, V, sigma, sweeps);
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

extern "C" int lb_svd_randn(float* out, long long n, unsigned long long seed, void* stream) {
  if (n <= 0) return LB_ERR_SHAPE;
  const int blocks = static_cast<int>((n + 255) / 256);
  randn_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, n, seed);
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

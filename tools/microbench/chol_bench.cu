// Microbenchmark of in-shared-memory Cholesky variants for the 6h x 6h Psi matrix (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo chol_bench.cu -o chol_bench
// Prints, per variant: cycles per factorisation for one CTA alone on the GPU (latency) and the
// aggregate factorisations/s with 8 CTAs resident per SM (throughput), plus the max error of L.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr unsigned kFull = 0xffffffffu;
__device__ long long g_phase[8];
#define PH_TIC() long long ph_a0_ = 0, ph_a1_ = 0, ph_a2_ = 0, ph_a3_ = 0; long long ph_t_ = clock64()
#define PH_TOC(k, thr) do { const long long now_ = clock64(); ph_a##k##_ += now_ - ph_t_; ph_t_ = now_; } while (0)
#define PH_END(thr) do { if (blockIdx.x == 0 && threadIdx.x == (thr)) { g_phase[0] += ph_a0_; g_phase[1] += ph_a1_; g_phase[2] += ph_a2_; g_phase[3] += ph_a3_; } } while (0)
__device__ __forceinline__ int prow(int i) { return i * (i + 1) / 2; }
__device__ __forceinline__ double2 ld2(const double* p) { return make_double2(p[0], p[1]); }

// ---------------------------------------------------------------- V0: the kernel's current routine (4-wide, hand pipelined)
template <int N6>
__device__ __noinline__ void chol_v0(double* psi, double* rdiag, double* blk44, int* flag) {
  const int i = threadIdx.x;
  const bool row_ok = i < N6;
  double* row_i = psi + prow(row_ok ? i : 0);
  if (i == 0) *flag = 0;
  PH_TIC();
#pragma unroll 1
  for (int j0 = 0; j0 < N6; j0 += 4) {
    PH_TOC(3, N6 - 1);
    const int w = N6 - j0 < 4 ? N6 - j0 : 4;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const bool in_play = row_ok && i >= j0;
    if (in_play) {
      const double* r2 = row_i;
      const double* p0 = psi + prow(j0);
      const double* p1 = psi + prow(j0 + (w > 1 ? 1 : 0));
      const double* p2 = psi + prow(j0 + (w > 2 ? 2 : 0));
      const double* p3 = psi + prow(j0 + (w > 3 ? 3 : 0));
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] = (c < w && j0 + c <= i) ? row_i[j0 + c] : 0.0;
      const int ng = j0 >> 1;
      double2 a = ld2(r2), b0 = ld2(p0), b1 = ld2(p1), b2 = ld2(p2), b3 = ld2(p3);
#pragma unroll 2
      for (int g = 0; g < ng; ++g) {
        const int gn = g + 1 < ng ? g + 1 : g;
        const double2 an = ld2(r2 + 2 * gn), c0 = ld2(p0 + 2 * gn), c1 = ld2(p1 + 2 * gn), c2 = ld2(p2 + 2 * gn), c3 = ld2(p3 + 2 * gn);
        acc[0] = fma(-a.x, b0.x, acc[0]); acc[1] = fma(-a.x, b1.x, acc[1]);
        acc[2] = fma(-a.x, b2.x, acc[2]); acc[3] = fma(-a.x, b3.x, acc[3]);
        acc[0] = fma(-a.y, b0.y, acc[0]); acc[1] = fma(-a.y, b1.y, acc[1]);
        acc[2] = fma(-a.y, b2.y, acc[2]); acc[3] = fma(-a.y, b3.y, acc[3]);
        a = an; b0 = c0; b1 = c1; b2 = c2; b3 = c3;
      }
      if (i < j0 + w) {
#pragma unroll
        for (int c = 0; c < 4; ++c) if (j0 + c <= i) blk44[4 * (i - j0) + c] = acc[c];
      }
    }
    PH_TOC(0, N6 - 1);
    __syncthreads();
    PH_TOC(1, N6 - 1);
    if (in_play) {
      double a[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r][c] = (r < w) ? blk44[4 * r + c] : (r == c ? 1.0 : 0.0);
      double rd[4];
      bool bad = false;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double d = a[c][c];
#pragma unroll
        for (int k = 0; k < c; ++k) d = fma(-a[c][k], a[c][k], d);
        if (!(d > 0.0)) { bad = true; d = 1e-300; }
        rd[c] = rsqrt(d);
#pragma unroll
        for (int r = c + 1; r < 4; ++r) {
          double v = a[r][c];
#pragma unroll
          for (int k = 0; k < c; ++k) v = fma(-a[r][k], a[c][k], v);
          a[r][c] = v * rd[c];
        }
      }
      if (i < j0 + w) {
        const int r = i - j0;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          if (rr == r) {
#pragma unroll
            for (int c = 0; c < rr; ++c) row_i[j0 + c] = a[rr][c];
            rdiag[i] = rd[rr];
          }
        }
        if (bad && r == 0) *flag = 1;
      } else {
        double x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double v = acc[c];
#pragma unroll
          for (int k = 0; k < c; ++k) v = fma(-x[k], a[c][k], v);
          x[c] = v * rd[c];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) if (c < w) row_i[j0 + c] = x[c];
      }
    }
    PH_TOC(2, N6 - 1);
    __syncthreads();
  }
  PH_END(N6 - 1);
}

// ---------------------------------------------------------------- V1: generic W-wide panel, compiler-scheduled loop
template <int N6, int W, int UNROLL>
__device__ __noinline__ void chol_panel(double* psi, double* rdiag, double* blk, int* flag) {
  const int i = threadIdx.x;
  const bool row_ok = i < N6;
  double* row_i = psi + prow(row_ok ? i : 0);
  if (i == 0) *flag = 0;
#pragma unroll 1
  for (int j0 = 0; j0 < N6; j0 += W) {
    const int w = N6 - j0 < W ? N6 - j0 : W;
    double acc[W];
    const bool in_play = row_ok && i >= j0;
    if (in_play) {
      const double* p[W];
#pragma unroll
      for (int c = 0; c < W; ++c) {
        p[c] = psi + prow(j0 + (c < w ? c : 0));
        acc[c] = (c < w && j0 + c <= i) ? row_i[j0 + c] : 0.0;
      }
#pragma unroll UNROLL
      for (int k = 0; k < j0; k += 2) {
        const double2 a = ld2(row_i + k);
#pragma unroll
        for (int c = 0; c < W; ++c) {
          const double2 b = ld2(p[c] + k);
          acc[c] = fma(-a.x, b.x, acc[c]);
          acc[c] = fma(-a.y, b.y, acc[c]);
        }
      }
      if (i < j0 + w) {
#pragma unroll
        for (int c = 0; c < W; ++c) if (j0 + c <= i) blk[W * (i - j0) + c] = acc[c];
      }
    }
    __syncthreads();
    if (in_play) {
      double a[W][W];
#pragma unroll
      for (int r = 0; r < W; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r][c] = (r < w) ? blk[W * r + c] : (r == c ? 1.0 : 0.0);
      double rd[W];
      bool bad = false;
#pragma unroll
      for (int c = 0; c < W; ++c) {
        double d = a[c][c];
        if (!(d > 0.0)) { bad = true; d = 1e-300; }
        rd[c] = rsqrt(d);
#pragma unroll
        for (int r = c + 1; r < W; ++r) a[r][c] *= rd[c];
        // right-looking inside the block: every later entry gets its update as soon as the column exists
#pragma unroll
        for (int r = c + 1; r < W; ++r)
#pragma unroll
          for (int cc = c + 1; cc <= r; ++cc) a[r][cc] = fma(-a[r][c], a[cc][c], a[r][cc]);
      }
      if (i < j0 + w) {
        const int r = i - j0;
#pragma unroll
        for (int rr = 0; rr < W; ++rr) {
          if (rr == r) {
#pragma unroll
            for (int c = 0; c < rr; ++c) row_i[j0 + c] = a[rr][c];
            rdiag[i] = rd[rr];
          }
        }
        if (bad && r == 0) *flag = 1;
      } else {
        double x[W];
#pragma unroll
        for (int c = 0; c < W; ++c) {
          double v = acc[c];
#pragma unroll
          for (int k = 0; k < c; ++k) v = fma(-x[k], a[c][k], v);
          x[c] = v * rd[c];
        }
#pragma unroll
        for (int c = 0; c < W; ++c) if (c < w) row_i[j0 + c] = x[c];
      }
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------- V6: FP64 tensor-core (DMMA m8n8k4) left-looking update,
// 8-wide panels; phase 2 (8x8 diagonal block in registers + row substitution) as in the scalar variants
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int N6>
__device__ __noinline__ void chol_dmma(double* psi, double* rdiag, double* blk, int* flag) {
  constexpr int W = 8;
  constexpr int NB = (N6 + 7) / 8;
  constexpr int NW = (N6 + 31) / 32;
  constexpr int TPW = (NB + NW - 1) / NW;
  const int i = threadIdx.x, lane = i & 31, wid = i >> 5, g = lane >> 2, q = lane & 3;
  const bool row_ok = i < N6;
  double* row_i = psi + prow(row_ok ? i : 0);
  if (i == 0) *flag = 0;
  PH_TIC();
#pragma unroll 1
  for (int J = 0; J < NB; ++J) {
    PH_TOC(3, N6 - 1);
    const int j0 = 8 * J;
    const int w = N6 - j0 < W ? N6 - j0 : W;
    {
      double d[TPW][2];
      const double* ap[TPW];
      const int brow = j0 + g < N6 ? j0 + g : N6 - 1;
      const double* bp = psi + prow(brow) + q;
#pragma unroll
      for (int s = 0; s < TPW; ++s) {
        const int arow = 8 * (J + wid + s * NW) + g;
        ap[s] = psi + prow(arow < N6 ? arow : N6 - 1) + q;
        d[s][0] = d[s][1] = 0.0;
      }
#pragma unroll 2
      for (int k0 = 0; k0 < j0; k0 += 4) {
        const double b = bp[k0];
#pragma unroll
        for (int s = 0; s < TPW; ++s) {
          if (J + wid + s * NW < NB) {     // warp-uniform
            const double a = ap[s][k0];
            dmma884(d[s][0], d[s][1], a, b);
          }
        }
      }
#pragma unroll
      for (int s = 0; s < TPW; ++s) {
        const int I = J + wid + s * NW;
        if (I < NB) {
          const int row = 8 * I + g;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = j0 + 2 * q + e;
            if (row < N6 && col <= row) {
              double* pe = psi + prow(row) + col;
              if (I == J) blk[8 * g + 2 * q + e] = *pe - d[s][e];
              else *pe -= d[s][e];
            }
          }
        }
      }
    }
    PH_TOC(0, N6 - 1);
    __syncthreads();
    PH_TOC(1, N6 - 1);
    const bool in_play = row_ok && i >= j0;
    if (in_play) {
      double a[W][W];
#pragma unroll
      for (int r = 0; r < W; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r][c] = (r < w) ? blk[W * r + c] : (r == c ? 1.0 : 0.0);
      double x[W];
      const bool panel_row = i < j0 + w;
#pragma unroll
      for (int c = 0; c < W; ++c) x[c] = (!panel_row && c < w) ? row_i[j0 + c] : 0.0;
      double rd[W];
      bool bad = false;
#pragma unroll
      for (int c = 0; c < W; ++c) {
        double dg = a[c][c];
        if (!(dg > 0.0)) { bad = true; dg = 1e-300; }
        rd[c] = rsqrt(dg);
#pragma unroll
        for (int r = c + 1; r < W; ++r) a[r][c] *= rd[c];
#pragma unroll
        for (int r = c + 1; r < W; ++r)
#pragma unroll
          for (int cc = c + 1; cc <= r; ++cc) a[r][cc] = fma(-a[r][c], a[cc][c], a[r][cc]);
        // the row below the block: x L_JJ^T = acc, eagerly (column c of the block dies here)
        x[c] *= rd[c];
#pragma unroll
        for (int r = c + 1; r < W; ++r) x[r] = fma(-x[c], a[r][c], x[r]);
      }
      if (panel_row) {
        const int r = i - j0;
#pragma unroll
        for (int rr = 0; rr < W; ++rr) {
          if (rr == r) {
#pragma unroll
            for (int c = 0; c < rr; ++c) row_i[j0 + c] = a[rr][c];
            rdiag[i] = rd[rr];
          }
        }
        if (bad && r == 0) *flag = 1;
      } else {
#pragma unroll
        for (int c = 0; c < W; ++c) if (c < w) row_i[j0 + c] = x[c];
      }
    }
    PH_TOC(2, N6 - 1);
    __syncthreads();
  }
  PH_END(N6 - 1);
}


// ---------------------------------------------------------------- V9: FP64 tensor-core (DMMA m8n8k4) update AND triangular solve.
// 8 x 8 tiles.  Per block column J:  (1) every warp updates its tiles (I, J), I >= J, with DMMA over the factored columns;
// (2) ONE warp factors the 8 x 8 diagonal tile cooperatively -- lane (g, q) holds columns q and q + 4 of row g of [S | I],
// eight elimination steps turn it into [L^T | L^-1] with shuffles only (nobody factors the block redundantly);
// (3) every warp multiplies its tiles below the diagonal by L^-T with two DMMAs per tile.  3 barriers per 8 columns.
template <int N6>
__device__ __noinline__ void chol_dmma2(double* psi, double* rdiag, double* blk, double* linv, int* flag) {
  constexpr int NB = (N6 + 7) / 8;
  constexpr int NW = (N6 + 31) / 32;
  constexpr int TPW = (NB + NW - 1) / NW;
  const int i = threadIdx.x, lane = i & 31, wid = i >> 5, g = lane >> 2, q = lane & 3;
  if (i == 0) *flag = 0;
#pragma unroll 1
  for (int J = 0; J < NB; ++J) {
    const int j0 = 8 * J;
    // ---- (1) left-looking update of block column J
    {
      double d[TPW][2];
      const double* ap[TPW];
      const int brow = j0 + g < N6 ? j0 + g : N6 - 1;
      const double* bp = psi + prow(brow) + q;
#pragma unroll
      for (int s = 0; s < TPW; ++s) {
        const int arow = 8 * (J + wid + s * NW) + g;
        ap[s] = psi + prow(arow < N6 ? arow : N6 - 1) + q;
        d[s][0] = d[s][1] = 0.0;
      }
#pragma unroll 2
      for (int k0 = 0; k0 < j0; k0 += 4) {
        const double b = bp[k0];
#pragma unroll
        for (int s = 0; s < TPW; ++s) {
          if (J + wid + s * NW < NB) {     // warp-uniform
            const double a = ap[s][k0];
            dmma884(d[s][0], d[s][1], a, b);
          }
        }
      }
#pragma unroll
      for (int s = 0; s < TPW; ++s) {
        const int I = J + wid + s * NW;
        if (I < NB) {
          const int row = 8 * I + g;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = j0 + 2 * q + e;
            if (I == J) {
              // diagonal tile: full symmetric 8 x 8 into the side buffer, identity beyond the matrix
              double v;
              if (row < N6 && col < N6) v = (col <= row ? psi[prow(row) + col] : psi[prow(col) + row]) - d[s][e];
              else v = (row == col) ? 1.0 : 0.0;
              blk[8 * g + 2 * q + e] = v;
            } else if (row < N6) {
              psi[prow(row) + col] -= d[s][e];
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- (2) diagonal tile: [S | I] -> [L^T | L^-1], one warp
    if (wid == 0) {
      double s0 = blk[8 * g + q], s1 = blk[8 * g + q + 4];
      double i0 = (g == q) ? 1.0 : 0.0, i1 = (g == q + 4) ? 1.0 : 0.0;
      bool bad = false;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double piv = __shfl_sync(kFull, (c < 4) ? s0 : s1, 4 * c + (c & 3));
        const double m = __shfl_sync(kFull, (c < 4) ? s0 : s1, (lane & ~3) | (c & 3));
        double dd = piv;
        if (!(dd > 0.0)) { bad = true; dd = 1e-300; }
        const double rd = rsqrt(dd);
        const double p0 = __shfl_sync(kFull, s0, 4 * c + q) * rd, p1 = __shfl_sync(kFull, s1, 4 * c + q) * rd;
        const double r0 = __shfl_sync(kFull, i0, 4 * c + q) * rd, r1 = __shfl_sync(kFull, i1, 4 * c + q) * rd;
        const double t = m * rd;
        if (g == c) { s0 = p0; s1 = p1; i0 = r0; i1 = r1; }
        else if (g > c) { s0 = fma(-t, p0, s0); s1 = fma(-t, p1, s1); i0 = fma(-t, r0, i0); i1 = fma(-t, r1, i1); }
      }
      // L^T[g][c] = L[c][g]: strictly lower entries of L go back into Psi, 1 / L[g][g] = L^-1[g][g] into rdiag
      if (q > g && j0 + q < N6) psi[prow(j0 + q) + j0 + g] = s0;
      if (q + 4 > g && j0 + q + 4 < N6) psi[prow(j0 + q + 4) + j0 + g] = s1;
      if (q == g && j0 + g < N6) rdiag[j0 + g] = i0;
      if (q + 4 == g && j0 + g < N6) rdiag[j0 + g] = i1;
      linv[8 * g + q] = i0;
      linv[8 * g + q + 4] = i1;
      if (bad && lane == 0) *flag = 1;
    }
    __syncthreads();
    // ---- (3) tiles below the diagonal: X = C' L^-T  (B fragment: B[k][n] = L^-1[n][k])
    {
      const double b0 = linv[8 * g + q], b1 = linv[8 * g + 4 + q];
#pragma unroll
      for (int s = 0; s < TPW; ++s) {
        const int I = J + wid + s * NW;
        if (I > J && I < NB) {
          const int row = 8 * I + g;
          double* rp = psi + prow(row < N6 ? row : N6 - 1) + j0;
          const double a0 = rp[q], a1 = rp[4 + q];
          double x0 = 0.0, x1 = 0.0;
          dmma884(x0, x1, a0, b0);
          dmma884(x0, x1, a1, b1);
          if (row < N6) { rp[2 * q] = x0; rp[2 * q + 1] = x1; }
        }
      }
    }
    __syncthreads();
  }
}

template <int N6, int VARIANT>
__global__ void __launch_bounds__(((N6 + 31) / 32) * 32, 8)
bench_kernel(const double* __restrict__ a_dense, double* __restrict__ l_out, long long* __restrict__ cycles, int reps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* psi = reinterpret_cast<double*>(smem_raw);
  double* rdiag = psi + N6 * (N6 + 1) / 2 + 64;
  double* blk = rdiag + N6;
  int* flag = reinterpret_cast<int*>(blk + 64);
  const int tid = threadIdx.x;
  long long total = 0;
  for (int rep = 0; rep < reps; ++rep) {
    if (tid < N6)
      for (int k = 0; k <= tid; ++k) psi[prow(tid) + k] = a_dense[tid * N6 + k];
    __syncthreads();
    const long long t0 = clock64();
    if (VARIANT == 0) chol_v0<N6>(psi, rdiag, blk, flag);
    if (VARIANT == 1) chol_panel<N6, 4, 2>(psi, rdiag, blk, flag);
    if (VARIANT == 2) chol_panel<N6, 6, 2>(psi, rdiag, blk, flag);
    if (VARIANT == 3) chol_panel<N6, 6, 1>(psi, rdiag, blk, flag);
    if (VARIANT == 4) chol_panel<N6, 8, 1>(psi, rdiag, blk, flag);
    if (VARIANT == 5) chol_panel<N6, 2, 4>(psi, rdiag, blk, flag);
    if (VARIANT == 6) chol_dmma<N6>(psi, rdiag, blk, flag);
    if (VARIANT == 9) chol_dmma2<N6>(psi, rdiag, blk, reinterpret_cast<double*>(flag) + 2, flag);
    total += clock64() - t0;
    __syncthreads();
  }
  if (blockIdx.x == 0) {
    if (tid == 0) cycles[0] = total / reps;
    if (tid < N6) {
      for (int k = 0; k < tid; ++k) l_out[tid * N6 + k] = psi[prow(tid) + k];
      l_out[tid * N6 + tid] = 1.0 / rdiag[tid];
    }
  }
}

template <int N6, int VARIANT>
void run(const char* name, const double* d_a, const std::vector<double>& l_ref, double* d_l, long long* d_cyc) {
  const int nt = ((N6 + 31) / 32) * 32;
  const size_t smem = (N6 * (N6 + 1) / 2 + 64 + N6 + 64 + 2) * sizeof(double) + (N6 == 60 ? 4600 : 1024);   // ~23 KB at N6=60, like the solver
  cudaFuncSetAttribute(bench_kernel<N6, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bench_kernel<N6, VARIANT>, cudaFuncAttributePreferredSharedMemoryCarveout, 90);
  const int reps = 20;
  bench_kernel<N6, VARIANT><<<1, nt, smem>>>(d_a, d_l, d_cyc, reps);
  long long zero[8] = {0}; cudaMemcpyToSymbol(g_phase, zero, sizeof(zero));
  bench_kernel<N6, VARIANT><<<1, nt, smem>>>(d_a, d_l, d_cyc, reps);
  cudaDeviceSynchronize();
  long long ph[8]; cudaMemcpyFromSymbol(ph, g_phase, sizeof(ph));
  if (ph[0] + ph[2] > 0) printf("      phases (last row, per factorisation): update %lld | barrier1 %lld | block+subst %lld | barrier2+loop %lld\n", ph[0] / reps, ph[1] / reps, ph[2] / reps, ph[3] / reps);
  long long solo = 0;
  cudaMemcpy(&solo, d_cyc, sizeof(solo), cudaMemcpyDeviceToHost);
  std::vector<double> l(N6 * N6);
  cudaMemcpy(l.data(), d_l, sizeof(double) * N6 * N6, cudaMemcpyDeviceToHost);
  double err = 0.0;
  for (int i = 0; i < N6; ++i) for (int k = 0; k <= i; ++k) err = fmax(err, fabs(l[i * N6 + k] - l_ref[i * N6 + k]));
  const int grid = 148 * 8 * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench_kernel<N6, VARIANT><<<grid, nt, smem>>>(d_a, d_l, d_cyc, reps);
  cudaEventRecord(e0);
  bench_kernel<N6, VARIANT><<<grid, nt, smem>>>(d_a, d_l, d_cyc, reps);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long loaded = 0;
  cudaMemcpy(&loaded, d_cyc, sizeof(loaded), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("N=%3d %-34s solo %7lld cyc | loaded (8 CTA/SM) %7lld cyc/CTA, %6.2f M factorisations/s (incl. reload) | max|L-Lref| %.2e %s\n",
         N6, name, solo, loaded, grid * (double)reps / ms * 1e-3, err, e == cudaSuccess ? "" : cudaGetErrorString(e));
}



// ---------------------------------------------------------------- V7: square layout, rows 16-byte aligned (stride LD, LD/2 odd):
// every shared-memory load of phase 1 is 128-bit (own row conflict-free per quarter warp, panel rows broadcast)
template <int N6, int W, int LD, int UNROLL>
__device__ __noinline__ void chol_sq(double* psi, double* rdiag, double* blk, int* flag) {
  const int i = threadIdx.x;
  const bool row_ok = i < N6;
  double* row_i = psi + (row_ok ? i : 0) * LD;
  if (i == 0) *flag = 0;
#pragma unroll 1
  for (int j0 = 0; j0 < N6; j0 += W) {
    const int w = N6 - j0 < W ? N6 - j0 : W;
    double acc[W];
    const bool in_play = row_ok && i >= j0;
    if (in_play) {
      const double* p[W];
#pragma unroll
      for (int c = 0; c < W; ++c) {
        p[c] = psi + (j0 + (c < w ? c : 0)) * LD;
        acc[c] = (c < w && j0 + c <= i) ? row_i[j0 + c] : 0.0;
      }
#pragma unroll UNROLL
      for (int k = 0; k < j0; k += 2) {
        const double2 a = *reinterpret_cast<const double2*>(row_i + k);
#pragma unroll
        for (int c = 0; c < W; ++c) {
          const double2 b = *reinterpret_cast<const double2*>(p[c] + k);
          acc[c] = fma(-a.x, b.x, acc[c]);
          acc[c] = fma(-a.y, b.y, acc[c]);
        }
      }
      if (i < j0 + w) {
#pragma unroll
        for (int c = 0; c < W; ++c) if (j0 + c <= i) blk[W * (i - j0) + c] = acc[c];
      }
    }
    __syncthreads();
    if (in_play) {
      double a[W][W];
#pragma unroll
      for (int r = 0; r < W; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r][c] = (r < w) ? blk[W * r + c] : (r == c ? 1.0 : 0.0);
      double rd[W];
      bool bad = false;
#pragma unroll
      for (int c = 0; c < W; ++c) {
        double d = a[c][c];
        if (!(d > 0.0)) { bad = true; d = 1e-300; }
        rd[c] = rsqrt(d);
#pragma unroll
        for (int r = c + 1; r < W; ++r) a[r][c] *= rd[c];
#pragma unroll
        for (int r = c + 1; r < W; ++r)
#pragma unroll
          for (int cc = c + 1; cc <= r; ++cc) a[r][cc] = fma(-a[r][c], a[cc][c], a[r][cc]);
      }
      if (i < j0 + w) {
        const int r = i - j0;
#pragma unroll
        for (int rr = 0; rr < W; ++rr) {
          if (rr == r) {
#pragma unroll
            for (int c = 0; c < rr; ++c) row_i[j0 + c] = a[rr][c];
            rdiag[i] = rd[rr];
          }
        }
        if (bad && r == 0) *flag = 1;
      } else {
        double x[W];
#pragma unroll
        for (int c = 0; c < W; ++c) {
          double v = acc[c];
#pragma unroll
          for (int k = 0; k < c; ++k) v = fma(-x[k], a[c][k], v);
          x[c] = v * rd[c];
        }
#pragma unroll
        for (int c = 0; c < W; ++c) if (c < w) row_i[j0 + c] = x[c];
      }
    }
    __syncthreads();
  }
}

template <int N6, int VARIANT>
__global__ void __launch_bounds__(((N6 + 31) / 32) * 32, 4)
bench_sq_kernel(const double* __restrict__ a_dense, double* __restrict__ l_out, long long* __restrict__ cycles, int reps) {
  constexpr int LD = N6 + 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* psi = reinterpret_cast<double*>(smem_raw);
  double* rdiag = psi + N6 * LD;
  double* blk = rdiag + N6;
  int* flag = reinterpret_cast<int*>(blk + 64);
  const int tid = threadIdx.x;
  long long total = 0;
  for (int rep = 0; rep < reps; ++rep) {
    if (tid < N6)
      for (int k = 0; k <= tid; ++k) psi[tid * LD + k] = a_dense[tid * N6 + k];
    __syncthreads();
    const long long t0 = clock64();
    if (VARIANT == 0) chol_sq<N6, 4, LD, 2>(psi, rdiag, blk, flag);
    if (VARIANT == 1) chol_sq<N6, 6, LD, 2>(psi, rdiag, blk, flag);
    if (VARIANT == 2) chol_sq<N6, 4, LD, 4>(psi, rdiag, blk, flag);
    total += clock64() - t0;
    __syncthreads();
  }
  if (blockIdx.x == 0) {
    if (tid == 0) cycles[0] = total / reps;
    if (tid < N6) {
      for (int k = 0; k < tid; ++k) l_out[tid * N6 + k] = psi[tid * LD + k];
      l_out[tid * N6 + tid] = 1.0 / rdiag[tid];
    }
  }
}

template <int N6, int VARIANT>
void run_sq(const char* name, const double* d_a, const std::vector<double>& l_ref, double* d_l, long long* d_cyc) {
  const int nt = ((N6 + 31) / 32) * 32;
  const size_t smem = (N6 * (N6 + 2) + N6 + 64 + 2) * sizeof(double);
  cudaFuncSetAttribute(bench_sq_kernel<N6, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bench_sq_kernel<N6, VARIANT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  const int reps = 20;
  bench_sq_kernel<N6, VARIANT><<<1, nt, smem>>>(d_a, d_l, d_cyc, reps);
  bench_sq_kernel<N6, VARIANT><<<1, nt, smem>>>(d_a, d_l, d_cyc, reps);
  cudaDeviceSynchronize();
  long long solo = 0;
  cudaMemcpy(&solo, d_cyc, sizeof(solo), cudaMemcpyDeviceToHost);
  std::vector<double> l(N6 * N6);
  cudaMemcpy(l.data(), d_l, sizeof(double) * N6 * N6, cudaMemcpyDeviceToHost);
  double err = 0.0;
  for (int i = 0; i < N6; ++i) for (int k = 0; k <= i; ++k) err = fmax(err, fabs(l[i * N6 + k] - l_ref[i * N6 + k]));
  const int grid = 148 * 7 * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench_sq_kernel<N6, VARIANT><<<grid, nt, smem>>>(d_a, d_l, d_cyc, reps);
  cudaEventRecord(e0);
  bench_sq_kernel<N6, VARIANT><<<grid, nt, smem>>>(d_a, d_l, d_cyc, reps);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long loaded = 0;
  cudaMemcpy(&loaded, d_cyc, sizeof(loaded), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("N=%3d %-34s solo %7lld cyc | loaded (smem-limited CTAs/SM) %7lld cyc/CTA, %6.2f M factorisations/s | max|L-Lref| %.2e %s\n",
         N6, name, solo, loaded, grid * (double)reps / ms * 1e-3, err, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// ================================================================ triangular solves  Psi x = b  (warp 0 only)
// T0: the kernel's current routine: lane owns rows lane, lane+32, ...; one pivot per step
template <int N6>
__device__ __noinline__ void tri_solve_t0(const double* psi, const double* rdiag, double* avec) {
  constexpr int RPL = (N6 + 31) / 32;
  const int lane = threadIdx.x;
  double x[RPL];
  int base[RPL];
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    const int i = lane + 32 * r;
    x[r] = i < N6 ? avec[i] : 0.0;
    base[r] = prow(i < N6 ? i : 0);
  }
#pragma unroll
  for (int slot = 0; slot < RPL; ++slot) {
    const int jend = N6 - 32 * slot < 32 ? N6 - 32 * slot : 32;
#pragma unroll 1
    for (int jj = 0; jj < jend; ++jj) {
      const int j = 32 * slot + jj;
      double xj = x[slot] * rdiag[j];
      xj = __shfl_sync(kFull, xj, jj);
      if (lane == jj) x[slot] = xj;
#pragma unroll
      for (int r = slot; r < RPL; ++r) {
        const int i = lane + 32 * r;
        if (i > j && i < N6) x[r] = fma(-psi[base[r] + j], xj, x[r]);
      }
    }
  }
#pragma unroll
  for (int slot = RPL - 1; slot >= 0; --slot) {
    const int jend = N6 - 32 * slot < 32 ? N6 - 32 * slot : 32;
#pragma unroll 1
    for (int jj = jend - 1; jj >= 0; --jj) {
      const int j = 32 * slot + jj;
      double xj = x[slot] * rdiag[j];
      xj = __shfl_sync(kFull, xj, jj);
      if (lane == jj) x[slot] = xj;
      const double* row_j = psi + prow(j);
#pragma unroll
      for (int r = 0; r <= slot; ++r) {
        const int i = lane + 32 * r;
        if (i < j) x[r] = fma(-row_j[i], xj, x[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RPL; ++r) { const int i = lane + 32 * r; if (i < N6) avec[i] = x[r]; }
}

// T3: T0 with the own diagonal reciprocals preloaded (no shared-memory load on the pivot chain)
template <int N6>
__device__ __noinline__ void tri_solve_t3(const double* psi, const double* rdiag, double* avec) {
  constexpr int RPL = (N6 + 31) / 32;
  const int lane = threadIdx.x;
  double x[RPL];
  int base[RPL];
  double rdo[RPL];
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    const int i = lane + 32 * r;
    x[r] = i < N6 ? avec[i] : 0.0;
    base[r] = prow(i < N6 ? i : 0);
    rdo[r] = rdiag[i < N6 ? i : 0];
  }
#pragma unroll
  for (int slot = 0; slot < RPL; ++slot) {
    const int jend = N6 - 32 * slot < 32 ? N6 - 32 * slot : 32;
#pragma unroll 1
    for (int jj = 0; jj < jend; ++jj) {
      const int j = 32 * slot + jj;
      double xj = x[slot] * rdo[slot];
      xj = __shfl_sync(kFull, xj, jj);
      if (lane == jj) x[slot] = xj;
#pragma unroll
      for (int r = slot; r < RPL; ++r) {
        const int i = lane + 32 * r;
        if (i > j && i < N6) x[r] = fma(-psi[base[r] + j], xj, x[r]);
      }
    }
  }
#pragma unroll
  for (int slot = RPL - 1; slot >= 0; --slot) {
    const int jend = N6 - 32 * slot < 32 ? N6 - 32 * slot : 32;
#pragma unroll 1
    for (int jj = jend - 1; jj >= 0; --jj) {
      const int j = 32 * slot + jj;
      double xj = x[slot] * rdo[slot];
      xj = __shfl_sync(kFull, xj, jj);
      if (lane == jj) x[slot] = xj;
      const double* row_j = psi + prow(j);
#pragma unroll
      for (int r = 0; r <= slot; ++r) {
        const int i = lane + 32 * r;
        if (i < j) x[r] = fma(-row_j[i], xj, x[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RPL; ++r) { const int i = lane + 32 * r; if (i < N6) avec[i] = x[r]; }
}


// T1: lane owns RPL CONSECUTIVE rows; a step eliminates the RPL pivots of one lane: local RPL x RPL
// triangular solve, RPL shuffles in flight together, RPL FMAs per owned row.
template <int N6, bool PREFETCH>
__device__ __noinline__ void tri_solve_t1(const double* psi, const double* rdiag, double* avec) {
  constexpr int RPL = (N6 + 31) / 32;
  constexpr int NL = N6 / RPL;
  static_assert(NL * RPL == N6, "rows must split evenly over the lanes");
  const int lane = threadIdx.x;
  const bool active = lane < NL;
  const int i0 = active ? lane * RPL : 0;
  double x[RPL], rd[RPL], lb[RPL][RPL];
  const double* rowp[RPL];
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    x[r] = active ? avec[i0 + r] : 0.0;
    rd[r] = rdiag[i0 + r];
    rowp[r] = psi + prow(i0 + r);
#pragma unroll
    for (int c = 0; c < r; ++c) lb[r][c] = rowp[r][i0 + c];
  }
  // ---- forward: L y = b
  double l[RPL][RPL];
  if (PREFETCH) {
#pragma unroll
    for (int r = 0; r < RPL; ++r)
#pragma unroll
      for (int c = 0; c < RPL; ++c) l[r][c] = rowp[r][c];      // pivot lane 0's columns
  }
#pragma unroll 1
  for (int p = 0; p < NL; ++p) {
    double ln[RPL][RPL];
    if (PREFETCH) {
      const int pn = p + 1 < NL ? p + 1 : p;
#pragma unroll
      for (int r = 0; r < RPL; ++r)
#pragma unroll
        for (int c = 0; c < RPL; ++c) ln[r][c] = rowp[r][pn * RPL + c];   // over-reads beyond the diagonal are never used
    } else {
#pragma unroll
      for (int r = 0; r < RPL; ++r)
#pragma unroll
        for (int c = 0; c < RPL; ++c) l[r][c] = rowp[r][p * RPL + c];
    }
    double y[RPL], yb[RPL];
#pragma unroll
    for (int c = 0; c < RPL; ++c) {
      double v = x[c];
#pragma unroll
      for (int cc = 0; cc < c; ++cc) v = fma(-lb[c][cc], y[cc], v);
      y[c] = v * rd[c];
    }
#pragma unroll
    for (int c = 0; c < RPL; ++c) yb[c] = __shfl_sync(kFull, y[c], p);
    {
      const bool own = lane == p, later = lane > p;
#pragma unroll
      for (int c = 0; c < RPL; ++c)
#pragma unroll
        for (int r = 0; r < RPL; ++r) x[r] = fma(later ? -l[r][c] : 0.0, yb[c], x[r]);
#pragma unroll
      for (int c = 0; c < RPL; ++c) x[c] = own ? y[c] : x[c];
    }
    if (PREFETCH) {
#pragma unroll
      for (int r = 0; r < RPL; ++r)
#pragma unroll
        for (int c = 0; c < RPL; ++c) l[r][c] = ln[r][c];
    }
  }
  // ---- backward: L^T x = y
#pragma unroll 1
  for (int p = NL - 1; p >= 0; --p) {
    // rows p*RPL + c of L at my columns i0 + r
#pragma unroll
    for (int c = 0; c < RPL; ++c) {
      const double* rp = psi + prow(p * RPL + c) + i0;
#pragma unroll
      for (int r = 0; r < RPL; ++r) l[c][r] = rp[r];
    }
    double z[RPL], zb[RPL];
#pragma unroll
    for (int c = RPL - 1; c >= 0; --c) {
      double v = x[c];
#pragma unroll
      for (int cc = c + 1; cc < RPL; ++cc) v = fma(-lb[cc][c], z[cc], v);
      z[c] = v * rd[c];
    }
#pragma unroll
    for (int c = 0; c < RPL; ++c) zb[c] = __shfl_sync(kFull, z[c], p);
    {
      const bool own = lane == p, earlier = lane < p;
#pragma unroll
      for (int c = RPL - 1; c >= 0; --c)
#pragma unroll
        for (int r = 0; r < RPL; ++r) x[r] = fma(earlier ? -l[c][r] : 0.0, zb[c], x[r]);
#pragma unroll
      for (int c = 0; c < RPL; ++c) x[c] = own ? z[c] : x[c];
    }
  }
  if (active) {
#pragma unroll
    for (int r = 0; r < RPL; ++r) avec[i0 + r] = x[r];
  }
}

// T4: T1's ownership, but the pivot lane's diagonal solve is folded into the update coefficients OFF the dependent
// chain: every lane loads the pivot block (broadcast), inverts it (RPL x RPL lower triangular) and multiplies its own
// RPL x RPL block of L with it while the previous step's chain is in flight.  The chain per step is then
// shuffle(raw x of the pivot lane) -> RPL FMAs, instead of (RPL mul + fma) -> shuffle -> RPL FMAs.
template <int RPL>
__device__ __forceinline__ void lower_inverse(const double (*lpp)[RPL], const double* rdp, double (*inv)[RPL]) {
#pragma unroll
  for (int c = 0; c < RPL; ++c) {
    inv[c][c] = rdp[c];
#pragma unroll
    for (int cc = c - 1; cc >= 0; --cc) {
      double s = 0.0;
#pragma unroll
      for (int k = cc; k < c; ++k) s = fma(lpp[c][k], inv[k][cc], s);
      inv[c][cc] = -rdp[c] * s;
    }
  }
}

template <int N6>
__device__ __noinline__ void tri_solve_t4(const double* psi, const double* rdiag, double* avec) {
  constexpr int RPL = (N6 + 31) / 32;
  constexpr int NL = N6 / RPL;
  static_assert(NL * RPL == N6, "rows must split evenly over the lanes");
  const int lane = threadIdx.x;
  const bool active = lane < NL;
  const int i0 = active ? lane * RPL : 0;
  double x[RPL];
  const double* rowp[RPL];
#pragma unroll
  for (int r = 0; r < RPL; ++r) { x[r] = active ? avec[i0 + r] : 0.0; rowp[r] = psi + prow(i0 + r); }
  // ---- forward: L y = b.   coefficients of step p: cf[r][c] = (L[i0+r][pR..] * inv(L_pp))[c]; the pivot lane itself gets inv(L_pp)
  double cf[RPL][RPL];
  auto fwd_coef = [&](int p, double (*out)[RPL]) {
    double lpp[RPL][RPL], rdp[RPL], inv[RPL][RPL], l[RPL][RPL];
#pragma unroll
    for (int c = 0; c < RPL; ++c) {
      rdp[c] = rdiag[p * RPL + c];
#pragma unroll
      for (int k = 0; k < c; ++k) lpp[c][k] = psi[prow(p * RPL + c) + p * RPL + k];
    }
#pragma unroll
    for (int r = 0; r < RPL; ++r)
#pragma unroll
      for (int c = 0; c < RPL; ++c) l[r][c] = rowp[r][p * RPL + c];   // lanes <= p read past their diagonal: unused
    lower_inverse<RPL>(lpp, rdp, inv);
    const bool own = lane == p, later = lane > p;
#pragma unroll
    for (int r = 0; r < RPL; ++r)
#pragma unroll
      for (int c = 0; c < RPL; ++c) {
        double s = 0.0;
#pragma unroll
        for (int k = c; k < RPL; ++k) s = fma(l[r][k], inv[k][c], s);
        // later lanes: x_r -= s * xp_c;  pivot lane: x_r = sum_c inv[r][c] xp_c  (written as x_r := 0 + ...);  earlier lanes: untouched
        out[r][c] = later ? -s : (own ? (c <= r ? inv[r][c] : 0.0) : 0.0);
      }
  };
  fwd_coef(0, cf);
#pragma unroll 1
  for (int p = 0; p < NL; ++p) {
    double cn[RPL][RPL];
    fwd_coef(p + 1 < NL ? p + 1 : p, cn);
    double xb[RPL];
#pragma unroll
    for (int c = 0; c < RPL; ++c) xb[c] = __shfl_sync(kFull, x[c], p);
    const bool own = lane == p;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
      double v = own ? 0.0 : x[r];
#pragma unroll
      for (int c = 0; c < RPL; ++c) v = fma(cf[r][c], xb[c], v);
      x[r] = v;
    }
#pragma unroll
    for (int r = 0; r < RPL; ++r)
#pragma unroll
      for (int c = 0; c < RPL; ++c) cf[r][c] = cn[r][c];
  }
  // ---- backward: L^T x = y.   x_p = inv(L_pp)^T yhat_p;  earlier lanes: x_r -= sum_c L[pR+c][i0+r] x_p[c] = sum_c' cb[r][c'] yhat_p[c']
  auto bwd_coef = [&](int p, double (*out)[RPL]) {
    double lpp[RPL][RPL], rdp[RPL], inv[RPL][RPL], l[RPL][RPL];
#pragma unroll
    for (int c = 0; c < RPL; ++c) {
      rdp[c] = rdiag[p * RPL + c];
#pragma unroll
      for (int k = 0; k < c; ++k) lpp[c][k] = psi[prow(p * RPL + c) + p * RPL + k];
    }
#pragma unroll
    for (int c = 0; c < RPL; ++c) {
      const double* rp = psi + prow(p * RPL + c) + i0;            // row of the pivot, my columns
#pragma unroll
      for (int r = 0; r < RPL; ++r) l[c][r] = rp[r];
    }
    lower_inverse<RPL>(lpp, rdp, inv);
    const bool own = lane == p, earlier = lane < p;
#pragma unroll
    for (int r = 0; r < RPL; ++r)
#pragma unroll
      for (int cc = 0; cc < RPL; ++cc) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c <= cc; ++c) s = fma(l[c][r], inv[cc][c], s);
        out[r][cc] = earlier ? -s : (own ? (cc >= r ? inv[cc][r] : 0.0) : 0.0);
      }
  };
  bwd_coef(NL - 1, cf);
#pragma unroll 1
  for (int p = NL - 1; p >= 0; --p) {
    double cn[RPL][RPL];
    bwd_coef(p > 0 ? p - 1 : 0, cn);
    double xb[RPL];
#pragma unroll
    for (int c = 0; c < RPL; ++c) xb[c] = __shfl_sync(kFull, x[c], p);
    const bool own = lane == p;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
      double v = own ? 0.0 : x[r];
#pragma unroll
      for (int c = 0; c < RPL; ++c) v = fma(cf[r][c], xb[c], v);
      x[r] = v;
    }
#pragma unroll
    for (int r = 0; r < RPL; ++r)
#pragma unroll
      for (int c = 0; c < RPL; ++c) cf[r][c] = cn[r][c];
  }
  if (active) {
#pragma unroll
    for (int r = 0; r < RPL; ++r) avec[i0 + r] = x[r];
  }
}

// T5: T1 with the instruction count per step cut down: predicated FMAs with negated operands instead of
// negate + select + FMA, no per-step fix-up of the pivot lane (every lane finishes with ONE local solve of its
// own rows after the sweep: its x stops changing at its own pivot step), row pointers advanced by additions.
template <int N6>
__device__ __noinline__ void tri_solve_t5(const double* psi, const double* rdiag, double* avec) {
  constexpr int RPL = (N6 + 31) / 32;
  constexpr int NL = N6 / RPL;
  static_assert(NL * RPL == N6, "rows must split evenly over the lanes");
  const int lane = threadIdx.x;
  const bool active = lane < NL;
  const int i0 = active ? lane * RPL : 0;
  double x[RPL], rd[RPL], lb[RPL][RPL];
  const double* rowp[RPL];
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    x[r] = active ? avec[i0 + r] : 0.0;
    rd[r] = rdiag[i0 + r];
    rowp[r] = psi + prow(i0 + r);
#pragma unroll
    for (int c = 0; c < r; ++c) lb[r][c] = rowp[r][i0 + c];
  }
  // ---- forward: L y = b
#pragma unroll 1
  for (int p = 0; p < NL - 1; ++p) {
    double l[RPL][RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r)
#pragma unroll
      for (int c = 0; c < RPL; ++c) l[r][c] = rowp[r][p * RPL + c];   // lanes <= p read past their diagonal: value unused
    double y[RPL], yb[RPL];
#pragma unroll
    for (int c = 0; c < RPL; ++c) {
      double v = x[c];
#pragma unroll
      for (int cc = 0; cc < c; ++cc) v = fma(-lb[c][cc], y[cc], v);
      y[c] = v * rd[c];
    }
#pragma unroll
    for (int c = 0; c < RPL; ++c) yb[c] = __shfl_sync(kFull, y[c], p);
    if (lane > p) {
#pragma unroll
      for (int c = 0; c < RPL; ++c)
#pragma unroll
        for (int r = 0; r < RPL; ++r) x[r] = fma(-l[r][c], yb[c], x[r]);
    }
  }
  // every lane: y of its own rows (its x is final since its own pivot step)
#pragma unroll
  for (int c = 0; c < RPL; ++c) {
    double v = x[c];
#pragma unroll
    for (int cc = 0; cc < c; ++cc) v = fma(-lb[c][cc], x[cc], v);
    x[c] = v * rd[c];
  }
  // ---- backward: L^T x = y
  const double* rp = psi + prow((NL - 1) * RPL) + i0;          // row of the pivot (first of the lane's rows), my columns
  int rlen = (NL - 1) * RPL;                                    // prow(i) - prow(i - 1) = i
#pragma unroll 1
  for (int p = NL - 1; p > 0; --p) {
    double l[RPL][RPL];
    {
      const double* q = rp;
#pragma unroll
      for (int c = 0; c < RPL; ++c) {
#pragma unroll
        for (int r = 0; r < RPL; ++r) l[c][r] = q[r];
        q += rlen + c + 1;                                       // next row of the packed triangle
      }
    }
    double z[RPL], zb[RPL];
#pragma unroll
    for (int c = RPL - 1; c >= 0; --c) {
      double v = x[c];
#pragma unroll
      for (int cc = c + 1; cc < RPL; ++cc) v = fma(-lb[cc][c], z[cc], v);
      z[c] = v * rd[c];
    }
#pragma unroll
    for (int c = 0; c < RPL; ++c) zb[c] = __shfl_sync(kFull, z[c], p);
    if (lane < p) {
#pragma unroll
      for (int c = RPL - 1; c >= 0; --c)
#pragma unroll
        for (int r = 0; r < RPL; ++r) x[r] = fma(-l[c][r], zb[c], x[r]);
    }
    // step back RPL rows: prow(i - RPL) = prow(i) - sum_{k=0}^{RPL-1} (i - k)
#pragma unroll
    for (int k = 0; k < RPL; ++k) { rp -= rlen; --rlen; }
  }
#pragma unroll
  for (int c = RPL - 1; c >= 0; --c) {
    double v = x[c];
#pragma unroll
    for (int cc = c + 1; cc < RPL; ++cc) v = fma(-lb[cc][c], x[cc], v);
    x[c] = v * rd[c];
  }
  if (active) {
#pragma unroll
    for (int r = 0; r < RPL; ++r) avec[i0 + r] = x[r];
  }
}

template <int N6, int VARIANT>
__global__ void __launch_bounds__(((N6 + 31) / 32) * 32, 8)
solve_kernel(const double* __restrict__ a_dense, const double* __restrict__ rhs, double* __restrict__ x_out, long long* __restrict__ cycles, int reps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* psi = reinterpret_cast<double*>(smem_raw);
  double* rdiag = psi + N6 * (N6 + 1) / 2 + 64;
  double* blk = rdiag + N6;
  int* flag = reinterpret_cast<int*>(blk + 64);
  double* avec = blk + 66;
  const int tid = threadIdx.x;
  if (tid < N6)
    for (int k = 0; k <= tid; ++k) psi[prow(tid) + k] = a_dense[tid * N6 + k];
  __syncthreads();
  chol_v0<N6>(psi, rdiag, blk, flag);
  long long total = 0;
  for (int rep = 0; rep < reps; ++rep) {
    if (tid < N6) avec[tid] = rhs[tid];
    __syncthreads();
    const long long t0 = clock64();
    if (tid < 32) {
      if (VARIANT == 0) tri_solve_t0<N6>(psi, rdiag, avec);
      if (VARIANT == 1) tri_solve_t1<N6, false>(psi, rdiag, avec);
      if (VARIANT == 2) tri_solve_t1<N6, true>(psi, rdiag, avec);
      if (VARIANT == 3) tri_solve_t3<N6>(psi, rdiag, avec);
      if (VARIANT == 4) tri_solve_t4<N6>(psi, rdiag, avec);
      if (VARIANT == 5) tri_solve_t5<N6>(psi, rdiag, avec);
    }
    __syncthreads();
    total += clock64() - t0;
  }
  if (blockIdx.x == 0) {
    if (tid == 0) cycles[0] = total / reps;
    if (tid < N6) x_out[tid] = avec[tid];
  }
}

template <int N6, int VARIANT>
void run_solve(const char* name, const double* d_a, const double* d_b, const std::vector<double>& x_ref, double* d_x, long long* d_cyc) {
  const int nt = ((N6 + 31) / 32) * 32;
  const size_t smem = (N6 * (N6 + 1) / 2 + 64 + N6 + 66 + N6 + 2) * sizeof(double) + (N6 == 60 ? 4000 : 0);
  cudaFuncSetAttribute(solve_kernel<N6, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(solve_kernel<N6, VARIANT>, cudaFuncAttributePreferredSharedMemoryCarveout, 90);
  const int reps = 20;
  solve_kernel<N6, VARIANT><<<1, nt, smem>>>(d_a, d_b, d_x, d_cyc, reps);
  solve_kernel<N6, VARIANT><<<1, nt, smem>>>(d_a, d_b, d_x, d_cyc, reps);
  cudaDeviceSynchronize();
  long long solo = 0;
  cudaMemcpy(&solo, d_cyc, sizeof(solo), cudaMemcpyDeviceToHost);
  std::vector<double> x(N6);
  cudaMemcpy(x.data(), d_x, sizeof(double) * N6, cudaMemcpyDeviceToHost);
  double err = 0.0, nrm = 0.0;
  for (int i = 0; i < N6; ++i) { err = fmax(err, fabs(x[i] - x_ref[i])); nrm = fmax(nrm, fabs(x_ref[i])); }
  const int grid = 148 * 8 * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  solve_kernel<N6, VARIANT><<<grid, nt, smem>>>(d_a, d_b, d_x, d_cyc, reps);
  cudaEventRecord(e0);
  solve_kernel<N6, VARIANT><<<grid, nt, smem>>>(d_a, d_b, d_x, d_cyc, reps);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long loaded = 0;
  cudaMemcpy(&loaded, d_cyc, sizeof(loaded), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("N=%3d solve %-34s solo %7lld cyc | loaded (8 CTA/SM) %7lld cyc/CTA, kernel %.3f ms (1 chol + %d solves per CTA) | rel err %.2e %s\n",
         N6, name, solo, loaded, ms, reps, err / nrm, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int N6>
void run_all() {
  std::vector<double> m(N6 * N6), a(N6 * N6, 0.0), l(N6 * N6, 0.0);
  srand(7);
  for (auto& v : m) v = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < N6; ++i) for (int j = 0; j < N6; ++j) { double s = 0; for (int k = 0; k < N6; ++k) s += m[i * N6 + k] * m[j * N6 + k]; a[i * N6 + j] = s + (i == j ? 1.0 : 0.0); }
  for (int j = 0; j < N6; ++j) {
    double d = a[j * N6 + j]; for (int k = 0; k < j; ++k) d -= l[j * N6 + k] * l[j * N6 + k];
    l[j * N6 + j] = sqrt(d);
    for (int i = j + 1; i < N6; ++i) { double v = a[i * N6 + j]; for (int k = 0; k < j; ++k) v -= l[i * N6 + k] * l[j * N6 + k]; l[i * N6 + j] = v / l[j * N6 + j]; }
  }
  double *d_a, *d_l; long long* d_cyc;
  cudaMalloc(&d_a, sizeof(double) * N6 * N6); cudaMalloc(&d_l, sizeof(double) * N6 * N6); cudaMalloc(&d_cyc, 64);
  cudaMemcpy(d_a, a.data(), sizeof(double) * N6 * N6, cudaMemcpyHostToDevice);
  if (!getenv("SOLVE_ONLY")) {
  run<N6, 0>("v0 current (4-wide, hand pipelined)", d_a, l, d_l, d_cyc);
  run<N6, 1>("panel W=4 unroll 2", d_a, l, d_l, d_cyc);
  run<N6, 2>("panel W=6 unroll 2", d_a, l, d_l, d_cyc);
  run<N6, 3>("panel W=6 unroll 1", d_a, l, d_l, d_cyc);
  run<N6, 4>("panel W=8 unroll 1", d_a, l, d_l, d_cyc);
  run<N6, 5>("panel W=2 unroll 4", d_a, l, d_l, d_cyc);
  run<N6, 6>("DMMA m8n8k4 update, W=8", d_a, l, d_l, d_cyc);
  run<N6, 9>("DMMA update + shuffle diag + DMMA trsm", d_a, l, d_l, d_cyc);
  run_sq<N6, 0>("square LD=N+2, LDS.128, W=4 u2", d_a, l, d_l, d_cyc);
  run_sq<N6, 1>("square LD=N+2, LDS.128, W=6 u2", d_a, l, d_l, d_cyc);
  run_sq<N6, 2>("square LD=N+2, LDS.128, W=4 u4", d_a, l, d_l, d_cyc);
  }
  if (getenv("CHOL_ONLY")) return;
  {
    std::vector<double> b(N6), y(N6), x(N6);
    for (auto& v : b) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < N6; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= l[i * N6 + k] * y[k]; y[i] = v / l[i * N6 + i]; }
    for (int i = N6 - 1; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < N6; ++k) v -= l[k * N6 + i] * x[k]; x[i] = v / l[i * N6 + i]; }
    double *d_b, *d_x; cudaMalloc(&d_b, sizeof(double) * N6); cudaMalloc(&d_x, sizeof(double) * N6);
    cudaMemcpy(d_b, b.data(), sizeof(double) * N6, cudaMemcpyHostToDevice);
    run_solve<N6, 0>("t0 current (strided rows)", d_a, d_b, x, d_x, d_cyc);
    run_solve<N6, 1>("t1 consecutive rows, lane blocks", d_a, d_b, x, d_x, d_cyc);
    run_solve<N6, 2>("t1 + prefetch (forward)", d_a, d_b, x, d_x, d_cyc);
    run_solve<N6, 3>("t3 = t0 + own rdiag in registers", d_a, d_b, x, d_x, d_cyc);
    run_solve<N6, 4>("t4 = t1, pivot solve folded into coefs", d_a, d_b, x, d_x, d_cyc);
    run_solve<N6, 5>("t5 = t1, predicated, no own fix-up", d_a, d_b, x, d_x, d_cyc);
    cudaFree(d_b); cudaFree(d_x);
  }
  cudaFree(d_a); cudaFree(d_l); cudaFree(d_cyc);
}

int main() {
  run_all<60>();
  run_all<30>();
  run_all<120>();

  return 0;
}

// HBM-bound batch kernels for sm_100a (compiled with -fmad=false so the arithmetic is the x86-64 reference's
// operation-for-operation; these kernels are bandwidth bound, so the un-fused multiplies are free):
//   K4 polyfit    /root/reference/mpc_to_line/src/helpers.h:24-44 (+ Eigen 3.3.3 HouseholderQR.h:256-287,350-369,
//                 Householder.h:64-131), polyeval helpers.h:13-19
//   K5 rollout    /root/reference/global_kinematic_model/solution/main.cpp:36-62
//   K6 batch I/O  the per-call unpack / pack of MPC.cpp:152-177 and :253-256, batched: [B][K] <-> [K][B]
// Compute kernels read and write field-major (SoA) arrays -- consecutive threads touch consecutive doubles -- or, for
// the host entry points, the reference's per-problem (AoS) rows directly: element k of problem b sits at
// p[b * sb + k * sk] with (sb, sk) = (1, B) or (K, 1).  A warp's K strided loads of K-double rows cover the same 32-byte
// sectors between them (L1 holds the 32 x 8K bytes in between), so the DRAM traffic is that of the coalesced layout and
// the separate transpose launches (K6) are not needed on those paths.
#include <float.h>

#include "kernels.h"

namespace b200mpc {

// ---------------------------------------------------------------------------------------------
// K6: tiled transpose through shared memory; both the global read and the global write are coalesced.
// Tile = TB problems x K fields, stored field-major in shared memory with pitch TB+1.
constexpr int kIoThreads = 128;

__global__ void __launch_bounds__(kIoThreads) aos_to_soa_kernel(const double* __restrict__ in, double* __restrict__ out, int B, int K, int TB) {
  extern __shared__ double tile[];
  const int pitch = TB + 1;
  const size_t b0 = (size_t)blockIdx.x * TB;
  const int nb = (int)(B - b0 < (size_t)TB ? B - b0 : (size_t)TB);
  for (int i = threadIdx.x; i < nb * K; i += kIoThreads) tile[(i % K) * pitch + (i / K)] = in[b0 * K + i];
  __syncthreads();
  for (int j = threadIdx.x; j < nb * K; j += kIoThreads) {
    const int k = j / nb, b = j % nb;
    out[(size_t)k * B + b0 + b] = tile[k * pitch + b];
  }
}

__global__ void __launch_bounds__(kIoThreads) soa_to_aos_kernel(const double* __restrict__ in, double* __restrict__ out, int B, int K, int TB) {
  extern __shared__ double tile[];
  const int pitch = TB + 1;
  const size_t b0 = (size_t)blockIdx.x * TB;
  const int nb = (int)(B - b0 < (size_t)TB ? B - b0 : (size_t)TB);
  for (int j = threadIdx.x; j < nb * K; j += kIoThreads) {
    const int k = j / nb, b = j % nb;
    tile[k * pitch + b] = in[(size_t)k * B + b0 + b];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb * K; i += kIoThreads) out[b0 * K + i] = tile[(i % K) * pitch + (i / K)];
}

// very wide records (K > 2048): plain element-wise transpose, coalesced on the field-major side only
__global__ void __launch_bounds__(256) transpose_wide_kernel(const double* __restrict__ in, double* __restrict__ out, int B, int K, int to_soa) {
  const size_t n = (size_t)B * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t k = i / B, b = i % B;   // i indexes the field-major array
    if (to_soa) out[i] = in[b * K + k]; else out[b * K + k] = in[i];
  }
}

static int tile_problems(int K) { return K <= 48 ? 128 : (K <= 400 ? 32 : 8); }

template <class Kern>
static cudaError_t launch_transpose(Kern kern, const double* in, double* out, int B, int K, int to_soa, cudaStream_t stream) {
  if (B <= 0 || K <= 0) return cudaSuccess;
  if (K > 2048) {
    transpose_wide_kernel<<<1184, 256, 0, stream>>>(in, out, B, K, to_soa);
    return cudaGetLastError();
  }
  const int TB = tile_problems(K);
  const size_t smem = (size_t)K * (TB + 1) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<(B + TB - 1) / TB, kIoThreads, smem, stream>>>(in, out, B, K, TB);
  return cudaGetLastError();
}

cudaError_t launch_aos_to_soa(const double* in, double* out, int B, int K, cudaStream_t stream) {
  return launch_transpose(aos_to_soa_kernel, in, out, B, K, 1, stream);
}
cudaError_t launch_soa_to_aos(const double* in, double* out, int B, int K, cudaStream_t stream) {
  return launch_transpose(soa_to_aos_kernel, in, out, B, K, 0, stream);
}

// ---------------------------------------------------------------------------------------------
// K4: one fit per thread; the m x n Vandermonde matrix lives in registers (MM, NN compile-time).
template <int MM, int NN>
__device__ __forceinline__ void polyfit_regs(const double* xs, const double* ys, double* out) {
  double A[NN][MM];   // column major like Eigen
#pragma unroll
  for (int j = 0; j < MM; ++j) A[0][j] = 1.0;
#pragma unroll
  for (int i = 0; i < NN - 1; ++i)
#pragma unroll
    for (int j = 0; j < MM; ++j) A[i + 1][j] = A[i][j] * xs[j];   // helpers.h:34-38
  double c[MM], h[NN], ib[NN];   // ib[k] = 1 / R(k,k) for the back substitution
#pragma unroll
  for (int j = 0; j < MM; ++j) c[j] = ys[j];
  constexpr int size = MM < NN ? MM : NN;
#pragma unroll
  for (int k = 0; k < size; ++k) {   // HouseholderQR.h:274-286
    const int rr = MM - k;
    double tail = 0.0;
#pragma unroll
    for (int i = 1; i < rr; ++i) tail += A[k][k + i] * A[k][k + i];
    const double c0 = A[k][k];
    double beta, tau;
    if (tail <= DBL_MIN) {   // Householder.h:79-84
      tau = 0.0; beta = c0;
      ib[k] = 1.0 / c0;
#pragma unroll
      for (int i = 1; i < rr; ++i) A[k][k + i] = 0.0;
    } else {
      // Eigen: beta = -sign(c0) sqrt(c0^2 + tail), tail /= (c0 - beta), tau = (beta - c0) / beta.  One rsqrt gives both
      // |beta| and 1/|beta|, one reciprocal replaces the element-wise divisions (results differ by <= 2 ulp)
      const double n2 = c0 * c0 + tail, rn = rsqrt(n2);
      double ibeta = rn;
      beta = n2 * rn;
      if (c0 >= 0.0) { beta = -beta; ibeta = -ibeta; }
      const double inv = 1.0 / (c0 - beta);
#pragma unroll
      for (int i = 1; i < rr; ++i) A[k][k + i] = A[k][k + i] * inv;
      tau = (beta - c0) * ibeta;
      ib[k] = ibeta;
    }
    h[k] = tau; A[k][k] = beta;
#pragma unroll
    for (int j = k + 1; j < NN; ++j) {   // applyHouseholderOnTheLeft, Householder.h:113-131
      if (rr == 1) { A[j][k] *= 1.0 - tau; continue; }
      if (tau == 0.0) continue;
      double t = 0.0;
#pragma unroll
      for (int i = 1; i < rr; ++i) t += A[k][k + i] * A[j][k + i];
      t += A[j][k];
      A[j][k] -= tau * t;
#pragma unroll
      for (int i = 1; i < rr; ++i) A[j][k + i] -= tau * A[k][k + i] * t;
    }
  }
#pragma unroll
  for (int k = 0; k < size; ++k) {   // Q^T y, HouseholderQR.h:358-362
    const int rr = MM - k;
    const double tau = h[k];
    if (rr == 1) { c[k] *= 1.0 - tau; continue; }
    if (tau == 0.0) continue;
    double t = 0.0;
#pragma unroll
    for (int i = 1; i < rr; ++i) t += A[k][k + i] * c[k + i];
    t += c[k];
    c[k] -= tau * t;
#pragma unroll
    for (int i = 1; i < rr; ++i) c[k + i] -= tau * A[k][k + i] * t;
  }
#pragma unroll
  for (int i = size - 1; i >= 0; --i) {   // back substitution on the top triangle
    double s = c[i];
#pragma unroll
    for (int j = i + 1; j < size; ++j) s -= A[j][i] * out[j];
    out[i] = s * ib[i];   // the reciprocal of the diagonal is already there from the reflector (<= 1 ulp from s / R(i,i))
  }
}

// (sb, sk) of a K-field array in either layout
struct Lay {
  size_t sb, sk;
  __host__ __device__ Lay(int aos, int B, int K) : sb(aos ? (size_t)K : 1), sk(aos ? 1 : (size_t)B) {}
  __host__ __device__ size_t operator()(int b, int k) const { return (size_t)b * sb + (size_t)k * sk; }
};

template <int MM, int NN>
__global__ void __launch_bounds__(128) polyfit_kernel(const double* __restrict__ xs, const double* __restrict__ ys, int B,
                                                      double* __restrict__ coeffs, int aos) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const Lay in(aos, B, MM), outl(aos, B, NN);
  double x[MM], y[MM], o[NN];
#pragma unroll
  for (int j = 0; j < MM; ++j) { x[j] = xs[in(b, j)]; y[j] = ys[in(b, j)]; }
  polyfit_regs<MM, NN>(x, y, o);
#pragma unroll
  for (int i = 0; i < NN; ++i) coeffs[outl(b, i)] = o[i];
}

// generic shapes (m <= 16, order <= 7): same algorithm on thread-local arrays with run-time sizes
__global__ void __launch_bounds__(128) polyfit_generic_kernel(const double* __restrict__ xs, const double* __restrict__ ys, int B,
                                                              int m, int n, double* __restrict__ coeffs, int aos) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const Lay in(aos, B, m), outl(aos, B, n);
  double A[8 * 16], c[16], h[8], out[8];
  for (int j = 0; j < m; ++j) { A[j] = 1.0; c[j] = ys[in(b, j)]; }
  for (int i = 0; i < n - 1; ++i)
    for (int j = 0; j < m; ++j) A[(i + 1) * m + j] = A[i * m + j] * xs[in(b, j)];
  const int size = m < n ? m : n;
  for (int k = 0; k < size; ++k) {
    const int rr = m - k;
    double* col = A + k + k * m;
    double tail = 0.0;
    for (int i = 1; i < rr; ++i) tail += col[i] * col[i];
    const double c0 = col[0];
    double beta, tau;
    if (tail <= DBL_MIN) {
      tau = 0.0; beta = c0;
      for (int i = 1; i < rr; ++i) col[i] = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tail);
      if (c0 >= 0.0) beta = -beta;
      for (int i = 1; i < rr; ++i) col[i] = col[i] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    h[k] = tau; col[0] = beta;
    for (int j = k + 1; j < n; ++j) {
      double* cj = A + k + j * m;
      if (rr == 1) { cj[0] *= 1.0 - tau; continue; }
      if (tau == 0.0) continue;
      double t = 0.0;
      for (int i = 1; i < rr; ++i) t += col[i] * cj[i];
      t += cj[0];
      cj[0] -= tau * t;
      for (int i = 1; i < rr; ++i) cj[i] -= tau * col[i] * t;
    }
  }
  for (int k = 0; k < size; ++k) {
    const int rr = m - k;
    const double* col = A + k + k * m;
    const double tau = h[k];
    if (rr == 1) { c[k] *= 1.0 - tau; continue; }
    if (tau == 0.0) continue;
    double t = 0.0;
    for (int i = 1; i < rr; ++i) t += col[i] * c[k + i];
    t += c[k];
    c[k] -= tau * t;
    for (int i = 1; i < rr; ++i) c[k + i] -= tau * col[i] * t;
  }
  for (int i = size - 1; i >= 0; --i) {
    double s = c[i];
    for (int j = i + 1; j < size; ++j) s -= A[i + j * m] * out[j];
    out[i] = s / A[i + i * m];
  }
  for (int i = 0; i < n; ++i) coeffs[outl(b, i)] = out[i];
}

template <int MM, int NN>
static cudaError_t polyfit_launch_t(const double* xs, const double* ys, int B, double* coeffs, cudaStream_t stream, int aos) {
  polyfit_kernel<MM, NN><<<(B + 127) / 128, 128, 0, stream>>>(xs, ys, B, coeffs, aos);
  return cudaGetLastError();
}

cudaError_t launch_polyfit(const double* xs, const double* ys, int B, int m, int order, double* coeffs,
                           cudaStream_t stream, int aos) {
  if (B <= 0) return cudaSuccess;
  const int n = order + 1;
#define PF_CASE(MM, NN) if (m == MM && n == NN) return polyfit_launch_t<MM, NN>(xs, ys, B, coeffs, stream, aos);
  PF_CASE(2, 2) PF_CASE(3, 2) PF_CASE(3, 3) PF_CASE(4, 2) PF_CASE(4, 3) PF_CASE(4, 4) PF_CASE(5, 3) PF_CASE(5, 4)
  PF_CASE(6, 2) PF_CASE(6, 3) PF_CASE(6, 4) PF_CASE(8, 4)
#undef PF_CASE
  polyfit_generic_kernel<<<(B + 127) / 128, 128, 0, stream>>>(xs, ys, B, m, n, coeffs, aos);
  return cudaGetLastError();
}

// helpers.h:13-19: result += coeffs[i] * pow(x, i), i ascending
__global__ void __launch_bounds__(256) polyeval_kernel(const double* __restrict__ coeffs, int ncoef, const double* __restrict__ x,
                                                       double* __restrict__ y, int B, int aos) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const Lay cl(aos, B, ncoef);
  const double xv = x[b];
  double r = 0.0, xp = 1.0;
  for (int i = 0; i < ncoef; ++i) {
    r += coeffs[cl(b, i)] * xp;
    xp *= xv;
  }
  y[b] = r;
}
cudaError_t launch_polyeval(const double* coeffs, int ncoef, const double* x, double* y, int B, cudaStream_t stream, int aos) {
  if (B <= 0) return cudaSuccess;
  polyeval_kernel<<<(B + 255) / 256, 256, 0, stream>>>(coeffs, ncoef, x, y, B, aos);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Roadmap front-end: the producer of (coeffs, cte, epsi) in front of MPC::Solve.  For every vehicle
// pose (x, y, psi, v) in the road's global frame: nearest centre-line point (the argmin over squared distances that
// mpc_to_line/src/custom_MPC.cpp:177-185 attempts inside the CppAD tape), the window of kWin consecutive points
// starting there, global -> vehicle frame (translate by the position, rotate by -psi), degree-3 polyfit (K4), and the
// MPC state in the vehicle frame: (0, 0, 0, v, cte = p(0), epsi = -atan(p'(0))) as solution/main.cpp:34-37 defines them.
// The centre line is staged in shared memory once per block.
constexpr int kWin = 6;
__global__ void __launch_bounds__(128) roadmap_reference_kernel(const double* __restrict__ pose4, int B, const double* __restrict__ wp_xy,
                                                                int n_wp, double* __restrict__ state6, double* __restrict__ coeffs, int aos) {
  extern __shared__ double wp[];   // [2][n_wp]
  for (int i = threadIdx.x; i < n_wp; i += blockDim.x) { wp[i] = wp_xy[2 * i]; wp[n_wp + i] = wp_xy[2 * i + 1]; }
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const Lay pl(aos, B, 4), sl(aos, B, 6), cfl(aos, B, 4);
  const double x = pose4[pl(b, 0)], y = pose4[pl(b, 1)], psi = pose4[pl(b, 2)], v = pose4[pl(b, 3)];
  int best = 0;
  double bd = 1e300;
  for (int i = 0; i < n_wp; ++i) {
    const double dx = x - wp[i], dy = y - wp[n_wp + i];
    const double d = dx * dx + dy * dy;
    if (d < bd) { bd = d; best = i; }   // first minimum, like index_sort's ind[0]
  }
  if (best > n_wp - kWin) best = n_wp - kWin;
  double sp, cp;
  sincos(psi, &sp, &cp);
  double lx[kWin], ly[kWin], c[4];
#pragma unroll
  for (int j = 0; j < kWin; ++j) {
    const double dx = wp[best + j] - x, dy = wp[n_wp + best + j] - y;
    lx[j] = cp * dx + sp * dy;
    ly[j] = cp * dy - sp * dx;
  }
  polyfit_regs<kWin, 4>(lx, ly, c);
#pragma unroll
  for (int i = 0; i < 4; ++i) coeffs[cfl(b, i)] = c[i];
  state6[sl(b, 0)] = 0.0; state6[sl(b, 1)] = 0.0; state6[sl(b, 2)] = 0.0; state6[sl(b, 3)] = v;
  state6[sl(b, 4)] = c[0];          // polyeval(coeffs, 0) - 0
  state6[sl(b, 5)] = -atan(c[1]);   // 0 - atan(p'(0))
}
cudaError_t launch_roadmap_reference(const double* pose4, int B, const double* wp_xy, int n_wp, double* state6, double* coeffs,
                                     cudaStream_t stream, int aos) {
  if (B <= 0) return cudaSuccess;
  const size_t smem = (size_t)2 * n_wp * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(roadmap_reference_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  roadmap_reference_kernel<<<(B + 127) / 128, 128, smem, stream>>>(pose4, B, wp_xy, n_wp, state6, coeffs, aos);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K5: H Euler steps of the bicycle model per vehicle, global_kinematic_model/solution/main.cpp:56-59
// (note the evaluation order v / Lf * delta * dt there).
__global__ void __launch_bounds__(256) rollout_kernel(const double* __restrict__ state4, const double* __restrict__ act, int B, int H,
                                                      double dt, double Lf, double* __restrict__ out, int aos) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const Lay sl(aos, B, 4), al(aos, B, 2 * H), ol(aos, B, 4 * H);
  double x = state4[sl(b, 0)], y = state4[sl(b, 1)], psi = state4[sl(b, 2)], v = state4[sl(b, 3)];
  for (int s = 0; s < H; ++s) {
    const double delta = act[al(b, 2 * s)], a = act[al(b, 2 * s + 1)];
    double sp, cp;
    sincos(psi, &sp, &cp);
    const double nx = x + v * cp * dt;
    const double ny = y + v * sp * dt;
    const double np = psi + v / Lf * delta * dt;
    const double nv = v + a * dt;
    x = nx; y = ny; psi = np; v = nv;
    out[ol(b, 4 * s + 0)] = x;
    out[ol(b, 4 * s + 1)] = y;
    out[ol(b, 4 * s + 2)] = psi;
    out[ol(b, 4 * s + 3)] = v;
  }
}
cudaError_t launch_rollout(const double* state4, const double* act, int B, int H, double dt, double Lf, double* out,
                           cudaStream_t stream, int aos) {
  if (B <= 0 || H <= 0) return cudaSuccess;
  rollout_kernel<<<(B + 255) / 256, 256, 0, stream>>>(state4, act, B, H, dt, Lf, out, aos);
  return cudaGetLastError();
}

}  // namespace b200mpc

// PFNLayer in TRAINING form behind the fused decoration (sm_100a).
//
// Replaces second/second/pytorch/models/pointpillars.py:51-65 (Linear(C_in -> units, no bias) ->
// BatchNorm1d with batch statistics -> ReLU -> max over the T slots of a pillar) and its autograd
// backward, without ever writing the (P,T,C_in) decorated tensor or the (P,T,units) activations
// (64.8 MB and 460 MB per 30,000-pillar sample in the reference).
//
// The Linear has no bias and the decoration zeroes padded slots, so y = W f is linear in f and
// the batch statistics of y over all N = P*T slots follow from the first and second moments of
// the decorated features alone:
//     mean_c   = W_c . S1 / N                     S1 = sum over live slots of f        (C_in)
//     E[y_c^2] = W_c^T M W_c / N                  M  = sum over live slots of f f^T    (C_in x C_in)
// (padded slots add zeros).  So:
//   forward   pillar_moments_kernel  -> S1, M in float64 (one entry per lane, DFMA per live slot)
//             host (PyTorch, 64 numbers): mean, var -> scale = gamma * invstd, shift = beta - mean * scale
//             lv_pillar_pfn (the inference kernel) with that scale / shift -> (P, units)
//   backward  pillar_pfn_bwd_kernel: recomputes the decoration and y per pillar, finds the slot that won
//             the max per channel, and accumulates in float64
//                 dbeta_c  = sum g            dgamma_c = sum g * yhat        A[c,k] = sum g * f[t*,k]
//             over the (pillar, channel) pairs whose maximum is positive (ReLU) - g = dL/dout.
//             host: dW = (gamma invstd) (A - dbeta/N S1 - dgamma/N * invstd (W M - mean S1)), the BatchNorm
//             backward written out with the same moments.
// Algorithmic bytes per pillar: T*16 + 20 read per pass (+ units*4 of g in the backward).
#include <algorithm>

#include <cuda_fp16.h>

#include "lv_common.cuh"
#include "lv_decorate.cuh"

#define PT_WARPS 8
#define PT_MAX_ENTRIES 96   // C_in + C_in (C_in + 1) / 2 <= 65 for C_in <= 10: three entries per lane

struct PtParams {
  const float* voxels;
  const int32_t* num;
  const int32_t* coors;
  int64_t P;
};

// ---------------------------------------------------------------- moments of the decorated features
// moments[e], e < C_in: S1; then the upper triangle of M row by row (i <= j).
__global__ void __launch_bounds__(PT_WARPS * 32) pillar_moments_kernel(PtParams p, DecoCfg d, double* __restrict__ moments) {
  extern __shared__ __align__(16) float stage[];   // [PT_WARPS][T * LV_PFN_STRIDE]
  __shared__ unsigned char ei[PT_MAX_ENTRIES], ej[PT_MAX_ENTRIES];
  __shared__ double red[PT_WARPS][PT_MAX_ENTRIES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CIN = d.C_out;
  const int n_entries = CIN + CIN * (CIN + 1) / 2;
  if (threadIdx.x == 0) {
    int e = 0;
    for (int i = 0; i < CIN; ++i, ++e) { ei[e] = (unsigned char)i; ej[e] = 255; }   // f_i * 1
    for (int i = 0; i < CIN; ++i)
      for (int j = i; j < CIN; ++j, ++e) { ei[e] = (unsigned char)i; ej[e] = (unsigned char)j; }
    for (; e < PT_MAX_ENTRIES; ++e) { ei[e] = 0; ej[e] = 255; }
  }
  __syncthreads();
  float* st = stage + warp * (d.T * LV_PFN_STRIDE);
  int mi[3], mj[3];
  double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int q = 0; q < 3; ++q) { mi[q] = ei[lane + 32 * q]; mj[q] = ej[lane + 32 * q]; }
  const int64_t warps_total = (int64_t)gridDim.x * PT_WARPS;
  for (int64_t pil = (int64_t)blockIdx.x * PT_WARPS + warp; pil < p.P; pil += warps_total) {
    const float4* v = reinterpret_cast<const float4*>(p.voxels) + pil * d.T;
    const int num = __ldg(p.num + pil);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (lane < d.T) a = lv_ld_stream_f4(v + lane);            // every slot, like the forward kernel: the mean runs over all T
    if (lane + 32 < d.T) b = lv_ld_stream_f4(v + lane + 32);
    const int4 co = __ldg(reinterpret_cast<const int4*>(p.coors) + pil);  // b, z, y, x
    const int live = lv_decorate_stage(a, b, num, co.z, co.w, d, st, lane, LV_PFN_STRIDE);
    for (int t = 0; t < live; ++t) {
      const float* f = st + t * LV_PFN_STRIDE;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (q == 2 && n_entries <= 64) break;   // C_in <= 9: 54 entries, two per lane
        const double fi = (double)f[mi[q]];
        const double fj = mj[q] == 255 ? 1.0 : (double)f[mj[q]];
        acc[q] = fma(fi, fj, acc[q]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) red[warp][lane + 32 * q] = acc[q];
  __syncthreads();
  for (int e = threadIdx.x; e < n_entries; e += PT_WARPS * 32) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < PT_WARPS; ++w) s += red[w][e];
    atomicAdd(moments + e, s);
  }
}

// ---------------------------------------------------------------- backward of relu(BN(W f)) max-pooled over T
struct PtBwd {
  const float* weight;   // (units, C_in)
  const float* scale;    // gamma * invstd
  const float* shift;    // beta - mean * scale
  const float* mean;     // batch mean of y
  const float* invstd;   // 1 / sqrt(var + eps)
  const float* grad_out; // (P, units)
  double* acc;           // (units, 2 + C_in): dbeta, dgamma, A[c, :]
  int units;
};

template <int CIN, int UJ>
__global__ void __launch_bounds__(PT_WARPS * 32) pillar_pfn_bwd_kernel(PtParams p, DecoCfg d, PtBwd c) {
  extern __shared__ __align__(16) float stage[];   // [PT_WARPS][T * LV_PFN_STRIDE]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* st = stage + warp * (d.T * LV_PFN_STRIDE);
  float w[CIN][UJ], sc[UJ], sh[UJ], mu[UJ], is[UJ];
  double dbeta[UJ], dgamma[UJ], A[UJ][CIN];
#pragma unroll
  for (int j = 0; j < UJ; ++j) {
    const int ch = lane + 32 * j;
    sc[j] = __ldg(c.scale + ch); sh[j] = __ldg(c.shift + ch);
    mu[j] = __ldg(c.mean + ch); is[j] = __ldg(c.invstd + ch);
    dbeta[j] = 0.0; dgamma[j] = 0.0;
#pragma unroll
    for (int k = 0; k < CIN; ++k) { w[k][j] = __ldg(c.weight + ch * CIN + k); A[j][k] = 0.0; }
  }
  const int64_t warps_total = (int64_t)gridDim.x * PT_WARPS;
  for (int64_t pil = (int64_t)blockIdx.x * PT_WARPS + warp; pil < p.P; pil += warps_total) {
    const float4* v = reinterpret_cast<const float4*>(p.voxels) + pil * d.T;
    const int num = __ldg(p.num + pil);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (lane < d.T) a = lv_ld_stream_f4(v + lane);
    if (lane + 32 < d.T) b = lv_ld_stream_f4(v + lane + 32);
    const int4 co = __ldg(reinterpret_cast<const int4*>(p.coors) + pil);
    float g[UJ];
#pragma unroll
    for (int j = 0; j < UJ; ++j) g[j] = __ldg(c.grad_out + pil * c.units + lane + 32 * j);
    const int live = lv_decorate_stage(a, b, num, co.z, co.w, d, st, lane, LV_PFN_STRIDE);
    // the slot that wins the max of every channel: a padded slot (y = 0) competes whenever live < T
    float best[UJ], ybest[UJ];
    int tbest[UJ];
#pragma unroll
    for (int j = 0; j < UJ; ++j) {
      best[j] = live < d.T ? sh[j] : -INFINITY;
      ybest[j] = 0.f;
      tbest[j] = -1;
    }
    for (int t = 0; t < live; ++t) {
      const float4* f4 = reinterpret_cast<const float4*>(st + t * LV_PFN_STRIDE);   // broadcast reads
      float f[LV_PFN_STRIDE];
      *reinterpret_cast<float4*>(f) = f4[0];
      *reinterpret_cast<float4*>(f + 4) = f4[1];
      if (CIN > 8) *reinterpret_cast<float4*>(f + 8) = f4[2];
#pragma unroll
      for (int j = 0; j < UJ; ++j) {
        float y = 0.f;
#pragma unroll
        for (int k = 0; k < CIN; ++k) y = fmaf(f[k], w[k][j], y);   // the forward kernel's chain (lv_pfn_warp)
        const float act = fmaf(y, sc[j], sh[j]);
        if (act > best[j]) { best[j] = act; ybest[j] = y; tbest[j] = t; }
      }
    }
#pragma unroll
    for (int j = 0; j < UJ; ++j) {
      if (!(best[j] > 0.f)) continue;                 // ReLU: a non-positive maximum passes no gradient
      const double gd = (double)g[j];
      dbeta[j] += gd;
      dgamma[j] = fma(gd, (double)((ybest[j] - mu[j]) * is[j]), dgamma[j]);
      if (tbest[j] >= 0) {
        const float* f = st + tbest[j] * LV_PFN_STRIDE;
#pragma unroll
        for (int k = 0; k < CIN; ++k) A[j][k] = fma(gd, (double)f[k], A[j][k]);
      }
    }
    __syncwarp();
  }
  // lane's channels are its own: reduce over the warps of the CTA through shared memory, then one atomic per value
  __syncthreads();
  double* red = reinterpret_cast<double*>(stage);     // [PT_WARPS][UJ*32][2 + CIN] needs <= the stage: checked by the host
  const int per = 2 + CIN;
#pragma unroll
  for (int j = 0; j < UJ; ++j) {
    double* r = red + ((size_t)warp * (UJ * 32) + lane + 32 * j) * per;
    r[0] = dbeta[j]; r[1] = dgamma[j];
#pragma unroll
    for (int k = 0; k < CIN; ++k) r[2 + k] = A[j][k];
  }
  __syncthreads();
  const int n_vals = UJ * 32 * per;
  for (int i = threadIdx.x; i < n_vals; i += PT_WARPS * 32) {
    double s = 0.0;
#pragma unroll
    for (int wq = 0; wq < PT_WARPS; ++wq) s += red[(size_t)wq * n_vals + i];
    if (s != 0.0) atomicAdd(c.acc + i, s);
  }
}

template <int CIN>
static void pt_bwd_launch(int units, unsigned blocks, size_t smem, cudaStream_t stream, const PtParams& p, const DecoCfg& d,
                          const PtBwd& c) {
  if (units == 32) {
    cudaFuncSetAttribute(pillar_pfn_bwd_kernel<CIN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pillar_pfn_bwd_kernel<CIN, 1><<<blocks, PT_WARPS * 32, smem, stream>>>(p, d, c);
  } else if (units == 64) {
    cudaFuncSetAttribute(pillar_pfn_bwd_kernel<CIN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pillar_pfn_bwd_kernel<CIN, 2><<<blocks, PT_WARPS * 32, smem, stream>>>(p, d, c);
  } else {
    cudaFuncSetAttribute(pillar_pfn_bwd_kernel<CIN, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pillar_pfn_bwd_kernel<CIN, 4><<<blocks, PT_WARPS * 32, smem, stream>>>(p, d, c);
  }
}

static int pt_check(const char* fn, lv_handle* h, int64_t n_pillars, int32_t max_points, int32_t num_features, int32_t variant,
                    int32_t with_distance, const void* d_voxels, const void* d_coors, int* c_out) {
  LV_REQUIRE(h != nullptr, "%s: null handle", fn);
  LV_REQUIRE(n_pillars >= 0 && max_points > 0 && max_points <= 64 && num_features == 4,
             "%s: needs 4 features per point and max_points <= 64 (got %d, %d)", fn, num_features, max_points);
  *c_out = lv_pillar_out_channels(num_features, variant, with_distance);
  LV_REQUIRE(*c_out >= 8 && *c_out <= 10, "%s: bad variant %d", fn, variant);
  LV_REQUIRE(n_pillars == 0 || ((reinterpret_cast<uintptr_t>(d_coors) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_voxels) & 15) == 0),
             "%s: voxels and coors must be 16-byte aligned", fn);
  return LV_OK;
}

extern "C" int lv_pillar_pfn_moments(lv_handle* h, const float* d_voxels, const int32_t* d_num_points, const int32_t* d_coors,
                                     int64_t n_pillars, int32_t max_points, int32_t num_features, float vx, float vy,
                                     float x_offset, float y_offset, int32_t variant, int32_t with_distance,
                                     double* d_moments, lv_stream stream_) {
  int c_out = 0;
  LV_CHECK(pt_check("lv_pillar_pfn_moments", h, n_pillars, max_points, num_features, variant, with_distance, d_voxels, d_coors, &c_out));
  LV_REQUIRE(d_moments != nullptr, "lv_pillar_pfn_moments: null output");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const int n_entries = c_out + c_out * (c_out + 1) / 2;
  LV_CHECK_CUDA(cudaMemsetAsync(d_moments, 0, sizeof(double) * n_entries, stream));
  if (n_pillars == 0) return LV_OK;
  LV_REQUIRE(d_voxels && d_num_points && d_coors, "lv_pillar_pfn_moments: null pointer");
  PtParams p{d_voxels, d_num_points, d_coors, n_pillars};
  DecoCfg d{vx, vy, x_offset, y_offset, variant, with_distance ? 1 : 0, max_points, c_out};
  const size_t smem = (size_t)PT_WARPS * max_points * LV_PFN_STRIDE * sizeof(float);
  const unsigned blocks = (unsigned)std::min<int64_t>((int64_t)h->num_sms * 4, lv_div_up(n_pillars, PT_WARPS));
  pillar_moments_kernel<<<blocks, PT_WARPS * 32, smem, stream>>>(p, d, d_moments);
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}

extern "C" int lv_pillar_pfn_backward(lv_handle* h, const float* d_voxels, const int32_t* d_num_points, const int32_t* d_coors,
                                      int64_t n_pillars, int32_t max_points, int32_t num_features, float vx, float vy,
                                      float x_offset, float y_offset, int32_t variant, int32_t with_distance,
                                      const float* d_weight, const float* d_scale, const float* d_shift, const float* d_mean,
                                      const float* d_invstd, int32_t units, const float* d_grad_out, double* d_acc,
                                      lv_stream stream_) {
  int c_out = 0;
  LV_CHECK(pt_check("lv_pillar_pfn_backward", h, n_pillars, max_points, num_features, variant, with_distance, d_voxels, d_coors, &c_out));
  LV_REQUIRE(units == 32 || units == 64 || units == 128, "lv_pillar_pfn_backward: units must be 32, 64 or 128, got %d", units);
  LV_REQUIRE(d_acc != nullptr, "lv_pillar_pfn_backward: null output");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  LV_CHECK_CUDA(cudaMemsetAsync(d_acc, 0, sizeof(double) * units * (2 + c_out), stream));
  if (n_pillars == 0) return LV_OK;
  LV_REQUIRE(d_voxels && d_num_points && d_coors && d_weight && d_scale && d_shift && d_mean && d_invstd && d_grad_out,
             "lv_pillar_pfn_backward: null pointer");
  PtParams p{d_voxels, d_num_points, d_coors, n_pillars};
  DecoCfg d{vx, vy, x_offset, y_offset, variant, with_distance ? 1 : 0, max_points, c_out};
  PtBwd c{d_weight, d_scale, d_shift, d_mean, d_invstd, d_grad_out, d_acc, units};
  // the stage doubles as the cross-warp reduction buffer of the epilogue
  size_t smem = (size_t)PT_WARPS * max_points * LV_PFN_STRIDE * sizeof(float);
  const size_t red = (size_t)PT_WARPS * units * (2 + c_out) * sizeof(double);
  if (red > smem) smem = red;
  const unsigned blocks = (unsigned)std::min<int64_t>((int64_t)h->num_sms * 2, lv_div_up(n_pillars, PT_WARPS));
  if (c_out == 8) pt_bwd_launch<8>(units, blocks, smem, stream, p, d, c);
  else if (c_out == 9) pt_bwd_launch<9>(units, blocks, smem, stream, p, d, c);
  else pt_bwd_launch<10>(units, blocks, smem, stream, p, d, c);
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}
// ---------------------------------------------------------------- the 64-number ends of the two passes, on the device
// moments -> batch statistics of y = W f over all N = P*T slots -> what lv_pillar_pfn and the backward kernel take.
// stats: float [5][units] = scale (gamma * invstd), shift (beta - mean * scale), mean, invstd, biased variance.
__global__ void pfn_train_stats_kernel(const double* __restrict__ mom, const float* __restrict__ w, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, int cin, int units, double N, double eps,
                                       float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= units) return;
  const double* M = mom + cin;            // upper triangle, row by row
  double mean = 0.0, ey2 = 0.0;
  int e = 0;
  for (int i = 0; i < cin; ++i) {
    const double wi = (double)w[c * cin + i];
    mean += wi * mom[i];
    for (int j = i; j < cin; ++j, ++e) ey2 += (i == j ? 1.0 : 2.0) * wi * (double)w[c * cin + j] * M[e];
  }
  mean /= N;
  ey2 /= N;
  double var = ey2 - mean * mean;
  if (var < 0.0) var = 0.0;
  const double invstd = 1.0 / sqrt(var + eps);
  const double scale = (double)gamma[c] * invstd;
  stats[c] = (float)scale;
  stats[units + c] = (float)((double)beta[c] - mean * scale);
  stats[2 * units + c] = (float)mean;
  stats[3 * units + c] = (float)invstd;
  stats[4 * units + c] = (float)var;
}

// acc (dbeta, dgamma, A) + moments -> dW, dgamma, dbeta: the BatchNorm backward over all N slots written with the
// moments (sum_live f = S1, sum_live yhat_c f = invstd_c (W M - mean S1)_c).  Statistics are recomputed in float64.
__global__ void pfn_train_finish_kernel(const double* __restrict__ acc, const double* __restrict__ mom, const float* __restrict__ w,
                                        const float* __restrict__ gamma, int cin, int units, double N, double eps,
                                        float* __restrict__ d_w, float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= units) return;
  const double* M = mom + cin;
  double mean = 0.0, ey2 = 0.0;
  int e = 0;
  for (int i = 0; i < cin; ++i) {
    const double wi = (double)w[c * cin + i];
    mean += wi * mom[i];
    for (int j = i; j < cin; ++j, ++e) ey2 += (i == j ? 1.0 : 2.0) * wi * (double)w[c * cin + j] * M[e];
  }
  mean /= N;
  double var = ey2 / N - mean * mean;
  if (var < 0.0) var = 0.0;
  const double invstd = 1.0 / sqrt(var + eps);
  const double* a = acc + (size_t)c * (2 + cin);
  const double dbeta = a[0], dgamma = a[1];
  d_beta[c] = (float)dbeta;
  d_gamma[c] = (float)dgamma;
  for (int k = 0; k < cin; ++k) {
    double wm = 0.0;                       // (W M)[c, k], M symmetric
    for (int l = 0; l < cin; ++l) {
      const int i = l < k ? l : k, j = l < k ? k : l;
      wm += (double)w[c * cin + l] * M[i * cin - i * (i - 1) / 2 + (j - i)];
    }
    const double yhat_f = invstd * (wm - mean * mom[k]);
    d_w[c * cin + k] = (float)((double)gamma[c] * invstd * (a[2 + k] - dbeta / N * mom[k] - dgamma / N * yhat_f));
  }
}

extern "C" int lv_pillar_pfn_train_forward(lv_handle* h, const float* d_voxels, const int32_t* d_num_points, const int32_t* d_coors,
                                           int64_t n_pillars, int32_t max_points, int32_t num_features, float vx, float vy,
                                           float x_offset, float y_offset, int32_t variant, int32_t with_distance,
                                           const float* d_weight, const float* d_gamma, const float* d_beta, double eps,
                                           int32_t units, float* d_out, float* d_stats, double* d_moments, lv_stream stream_) {
  int c_out = 0;
  LV_CHECK(pt_check("lv_pillar_pfn_train_forward", h, n_pillars, max_points, num_features, variant, with_distance, d_voxels, d_coors, &c_out));
  LV_REQUIRE(units == 32 || units == 64 || units == 128, "lv_pillar_pfn_train_forward: units must be 32, 64 or 128, got %d", units);
  LV_REQUIRE(n_pillars * (int64_t)max_points >= 2, "lv_pillar_pfn_train_forward: batch statistics need at least two slots");
  LV_REQUIRE(d_weight && d_gamma && d_beta && d_out && d_stats && d_moments, "lv_pillar_pfn_train_forward: null pointer");
  LV_CHECK(lv_pillar_pfn_moments(h, d_voxels, d_num_points, d_coors, n_pillars, max_points, num_features, vx, vy, x_offset, y_offset,
                                 variant, with_distance, d_moments, stream_));
  cudaStream_t stream = (cudaStream_t)stream_;
  pfn_train_stats_kernel<<<1, 128, 0, stream>>>(d_moments, d_weight, d_gamma, d_beta, c_out, units,
                                                (double)n_pillars * (double)max_points, eps, d_stats);
  LV_LAUNCH_CHECK(h);
  return lv_pillar_pfn(h, d_voxels, d_num_points, d_coors, n_pillars, max_points, num_features, vx, vy, x_offset, y_offset, variant,
                       with_distance, d_weight, d_stats, d_stats + units, units, d_out, stream_);
}

extern "C" int lv_pillar_pfn_train_backward(lv_handle* h, const float* d_voxels, const int32_t* d_num_points, const int32_t* d_coors,
                                            int64_t n_pillars, int32_t max_points, int32_t num_features, float vx, float vy,
                                            float x_offset, float y_offset, int32_t variant, int32_t with_distance,
                                            const float* d_weight, const float* d_gamma, double eps, const float* d_stats,
                                            const double* d_moments, int32_t units, const float* d_grad_out, float* d_dweight,
                                            float* d_dgamma, float* d_dbeta, lv_stream stream_) {
  int c_out = 0;
  LV_CHECK(pt_check("lv_pillar_pfn_train_backward", h, n_pillars, max_points, num_features, variant, with_distance, d_voxels, d_coors, &c_out));
  LV_REQUIRE(units == 32 || units == 64 || units == 128, "lv_pillar_pfn_train_backward: units must be 32, 64 or 128, got %d", units);
  LV_REQUIRE(d_weight && d_gamma && d_stats && d_moments && d_grad_out && d_dweight && d_dgamma && d_dbeta,
             "lv_pillar_pfn_train_backward: null pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  LV_CHECK(h->pfn_acc.ensure(sizeof(double) * 128 * (2 + 10), stream));
  LV_CHECK(lv_pillar_pfn_backward(h, d_voxels, d_num_points, d_coors, n_pillars, max_points, num_features, vx, vy, x_offset, y_offset,
                                  variant, with_distance, d_weight, d_stats, d_stats + units, d_stats + 2 * units, d_stats + 3 * units,
                                  units, d_grad_out, h->pfn_acc.as<double>(), stream_));
  pfn_train_finish_kernel<<<1, 128, 0, stream>>>(h->pfn_acc.as<double>(), d_moments, d_weight, d_gamma, c_out, units,
                                                 (double)n_pillars * (double)max_points, eps, d_dweight, d_dgamma, d_dbeta);
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}


// ---------------------------------------------------------------- backward of PointPillarsScatter
// d feats[p, c] = d canvas[b, c, y, x] for coords[p] = (b, ., y, x) (pointpillars.py:444-476 under autograd: the
// transpose of `canvas[:, indices] = voxels`).  One warp per pillar, lane = channel: the reads are one 4-byte element
// per channel plane (that is what the layout gives), the writes one coalesced row.  T = float or __half.
template <typename T>
__global__ void __launch_bounds__(256) pillar_gather_kernel(const T* __restrict__ grad, const int32_t* __restrict__ coords,
                                                           int64_t P, int C, int B, int ny, int nx, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * 8;
  const int64_t plane = (int64_t)ny * nx;
  for (int64_t pil = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); pil < P; pil += warps) {
    const int4 co = __ldg(reinterpret_cast<const int4*>(coords) + pil);   // b, z, y, x
    const bool ok = co.x >= 0 && co.x < B && co.z >= 0 && co.z < ny && co.w >= 0 && co.w < nx;
    const T* src = grad + ((int64_t)co.x * C) * plane + (int64_t)co.z * nx + co.w;
    for (int c = lane; c < C; c += 32) out[pil * C + c] = ok ? src[(int64_t)c * plane] : T(0);
  }
}

extern "C" int lv_pillar_scatter_backward(lv_handle* h, const void* d_grad_canvas, const int32_t* d_coords, int64_t n_pillars,
                                          int32_t channels, int32_t batch_size, int32_t ny, int32_t nx, int32_t half,
                                          void* d_grad_feats, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_pillar_scatter_backward: null handle");
  LV_REQUIRE(n_pillars >= 0 && channels > 0 && batch_size >= 0 && ny > 0 && nx > 0, "lv_pillar_scatter_backward: bad sizes");
  if (n_pillars == 0) return LV_OK;
  LV_REQUIRE(d_grad_canvas && d_coords && d_grad_feats, "lv_pillar_scatter_backward: null pointer");
  LV_REQUIRE((reinterpret_cast<uintptr_t>(d_coords) & 15) == 0, "lv_pillar_scatter_backward: coords must be 16-byte aligned");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const unsigned blocks = (unsigned)std::min<int64_t>((int64_t)h->num_sms * 16, lv_div_up(n_pillars, 8));
  if (half)
    pillar_gather_kernel<__half><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __half*>(d_grad_canvas), d_coords, n_pillars,
                                                              channels, batch_size, ny, nx, reinterpret_cast<__half*>(d_grad_feats));
  else
    pillar_gather_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(d_grad_canvas), d_coords, n_pillars,
                                                             channels, batch_size, ny, nx, reinterpret_cast<float*>(d_grad_feats));
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}

// Conditioning front-end of the reference's "ax" models, fp32 CUDA-core kernels on the reference's
// channels-first [B, C, T] layout.  Everything here runs once per infer call at mel-frame or group-step rate:
//   k_cn_conv      Conv1d (model-level cond_layers / res_conv, efficient_model_ax.py:74-113,293-307) and, per output
//                  phase, ConvTranspose1d (TransposedUpsampleNet, glow_ax.py:201-242) as one implicit GEMM:
//                  rows = (utterance, q), y[b][n][q*out_stride + out_off] = act(sum_{ci,j} x[b][ci][q + tap_step*j + in_off]
//                  * w[n][ci][j] + bias[n]) * out_scale (+ res)
//   k_cn_resample  F.interpolate (linear align_corners True/False, nearest) with crop / accumulate
//                  (efficient_model_ax.py:171-182, glow_ax.py:228-241)
//   k_cn_deemph    inverse pre-emphasis y[n] = x[n] + coef*y[n-1] (scipy.signal.lfilter on the host in the
//                  reference, efficient_model_ax.py:351-355) and the inverse perceived-volume map (:343-345)
#include "cwg_common.cuh"

namespace cwg {
namespace {

constexpr int CBM = 64, CBN = 64, CBK = 16;

struct ConvP {
  int B, Cin, Tin, Cout, Tout, Kt;    // Kt = taps in w ([Cout][Cin][Kt])
  int Q;                              // rows per utterance
  int tap_step, in_off, out_stride, out_off, pad_mode;
  int act; float slope, out_scale;
  const float *x, *w, *bias, *res;
  float* y;
  long long w_phase_stride; int out_off_phase_step;   // blockIdx.z = phase (ConvTranspose1d)
};

// pad_mode: 0 zeros, 1 replicate, 2 reflect, 3 circular (nn.Conv1d padding_mode)
__device__ __forceinline__ int map_index(int i, int n, int mode, bool* ok) {
  *ok = true;
  if (i >= 0 && i < n) return i;
  switch (mode) {
    case 1: return i < 0 ? 0 : n - 1;
    case 2: { if (n == 1) return 0; int p = 2 * (n - 1); i %= p; if (i < 0) i += p; return i < n ? i : p - i; }
    case 3: { i %= n; return i < 0 ? i + n : i; }
    default: *ok = false; return 0;
  }
}

__device__ __forceinline__ float activate(float v, int act, float slope) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return v >= 0.f ? v : v * slope;
    case 3: return tanhf(v);
    case 4: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

__global__ void __launch_bounds__(256) k_cn_conv(ConvP p) {
  __shared__ float As[CBK][CBM + 4];
  __shared__ float Bs[CBK][CBN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const long long M = (long long)p.B * p.Q;
  const long long m0 = (long long)blockIdx.x * CBM;
  const int n0 = blockIdx.y * CBN;
  const int phase = blockIdx.z;
  const float* w = p.w + (size_t)phase * p.w_phase_stride;
  const int out_off = p.out_off + phase * p.out_off_phase_step;
  const int KD = p.Cin * p.Kt;
  float acc[4][4] = {};
  // loader: thread -> row (tid % 64) x 4 k's for A (coalesced along time), n (tid / 4) x 4 k's for B
  const int a_row = tid % CBM, a_k = tid / CBM;           // a_k in 0..3, +4 per step
  long long am = m0 + a_row;
  int ab = 0, aq = 0;
  const bool a_ok = am < M;
  if (a_ok) { ab = (int)(am / p.Q); aq = (int)(am - (long long)ab * p.Q); }
  for (int k0 = 0; k0 < KD; k0 += CBK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = a_k + 4 * r, kd = k0 + kk;
      float v = 0.f;
      if (a_ok && kd < KD) {
        const int ci = kd / p.Kt, j = kd - ci * p.Kt;
        bool ok; const int ti = map_index(aq + p.tap_step * j + p.in_off, p.Tin, p.pad_mode, &ok);
        if (ok) v = __ldg(p.x + ((size_t)ab * p.Cin + ci) * p.Tin + ti);
      }
      As[kk][a_row] = v;
      const int idx = tid + r * 256, n = idx / CBK, bk = idx % CBK;
      Bs[bk][n] = (n0 + n < p.Cout && k0 + bk < KD) ? __ldg(w + (size_t)(n0 + n) * KD + k0 + bk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][tx + 16 * i]; b[i] = Bs[kk][ty * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + tx + 16 * i;
    if (m >= M) continue;
    const int b = (int)(m / p.Q), q = (int)(m - (long long)b * p.Q);
    const int to = q * p.out_stride + out_off;
    if (to < 0 || to >= p.Tout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + ty * 4 + j;
      if (n >= p.Cout) continue;
      const size_t o = ((size_t)b * p.Cout + n) * p.Tout + to;
      float v = acc[i][j] + (p.bias ? __ldg(p.bias + n) : 0.f);
      v = activate(v, p.act, p.slope) * p.out_scale;
      if (p.res) v += __ldg(p.res + o);
      p.y[o] = v;
    }
  }
}

// y[b][c][t] (t < Tout) = (accumulate ? y : 0) + interp(x[b][c][:])[t + crop], interp length Tvirt
__global__ void k_cn_resample(const float* __restrict__ x, int B, int C, int Tin, long long x_bstride,
                              float* __restrict__ y, int Tout, long long y_bstride, int mode, int Tvirt, int crop,
                              float inv_scale, int accumulate) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * Tout) return;
  const int t = (int)(i % Tout); const long long bc = i / Tout;
  const int c = (int)(bc % C), b = (int)(bc / C);
  const float* xr = x + (size_t)b * x_bstride + (size_t)c * Tin;
  const int tv = t + crop;
  float v;
  if (mode == 0) {                                        // nearest: floor(dst * scale), scale = 1/scale_factor or in/out
    const float sc = inv_scale > 0.f ? inv_scale : (float)Tin / (float)Tvirt;
    v = xr[min((int)floorf((float)tv * sc), Tin - 1)];
  } else {
    float src;
    if (mode == 1) src = Tvirt > 1 ? (float)tv * ((float)(Tin - 1) / (float)(Tvirt - 1)) : 0.f;   // align_corners=True
    else { const float sc = inv_scale > 0.f ? inv_scale : (float)Tin / (float)Tvirt; src = fmaxf(((float)tv + 0.5f) * sc - 0.5f, 0.f); }
    const int i0 = min((int)src, Tin - 1), i1 = min(i0 + 1, Tin - 1);
    const float l1 = src - (float)i0;
    v = (1.f - l1) * xr[i0] + l1 * xr[i1];
  }
  float* yo = y + (size_t)b * y_bstride + (size_t)c * Tout + t;
  *yo = accumulate ? *yo + v : v;
}

// One block per utterance: chunked first-order recurrence in fp64 (the reference filters in fp64 on the host).
constexpr int DE_THREADS = 1024;
__global__ void __launch_bounds__(DE_THREADS) k_cn_deemph(const float* __restrict__ x, float* __restrict__ y, int T,
                                                         double coef, int vol_scaling) {
  __shared__ double s_end[DE_THREADS];
  __shared__ double s_carry[DE_THREADS];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* xr = x + (size_t)b * T; float* yr = y + (size_t)b * T;
  const int len = (T + DE_THREADS - 1) / DE_THREADS;
  const int t0 = min(tid * len, T), t1 = min(t0 + len, T);
  auto load = [&](int t) -> double {
    float v = xr[t];
    if (vol_scaling) {                                    // z>0: 10**log2(z); z<0: -(10**log2(-z))
      if (v > 0.f) v = exp10f(log2f(v)); else if (v < 0.f) v = -exp10f(log2f(-v));
    }
    return (double)v;
  };
  double st = 0.0;
  if (coef != 0.0) for (int t = t0; t < t1; ++t) st = load(t) + coef * st;
  s_end[tid] = st;
  __syncthreads();
  if (tid == 0) {                                         // carry into each chunk; a^len computed once
    double al = 1.0; for (int i = 0; i < len; ++i) al *= coef;
    double carry = 0.0;
    for (int i = 0; i < DE_THREADS; ++i) { s_carry[i] = carry; carry = s_end[i] + al * carry; }
  }
  __syncthreads();
  st = s_carry[tid];
  for (int t = t0; t < t1; ++t) { st = load(t) + coef * st; yr[t] = (float)st; }
}

}  // namespace
}  // namespace cwg

using namespace cwg;

extern "C" {

int cwg_conv1d(const float* x, int batch, int c_in, int t_in, const float* w, const float* bias, int c_out, int k,
               int padding, int pad_mode, int act, float slope, float out_scale, const float* res, float* y,
               void* cuda_stream) {
  CWG_REQUIRE(batch >= 1 && c_in >= 1 && c_out >= 1 && k >= 1 && padding >= 0, "bad conv1d shape");
  const int t_out = t_in + 2 * padding - (k - 1);
  CWG_REQUIRE(t_in >= 1 && t_out >= 1, "conv1d: empty output");
  CWG_REQUIRE(pad_mode >= 0 && pad_mode <= 3 && act >= 0 && act <= 4, "conv1d: bad pad_mode / act");
  CWG_REQUIRE(pad_mode != 2 || padding < t_in, "conv1d: reflect padding needs padding < T");
  CWG_REQUIRE(x && w && y, "NULL argument");
  ConvP p{};
  p.B = batch; p.Cin = c_in; p.Tin = t_in; p.Cout = c_out; p.Tout = t_out; p.Kt = k; p.Q = t_out;
  p.tap_step = 1; p.in_off = -padding; p.out_stride = 1; p.out_off = 0; p.pad_mode = pad_mode;
  p.act = act; p.slope = slope; p.out_scale = out_scale; p.x = x; p.w = w; p.bias = bias; p.res = res; p.y = y;
  const long long M = (long long)batch * p.Q;
  dim3 grid((unsigned)((M + CBM - 1) / CBM), (c_out + CBN - 1) / CBN, 1);
  k_cn_conv<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(p);
  CWG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int cwg_conv_transpose1d(const float* x, int batch, int c_in, int t_in, const float* w_phases, const float* bias,
                         int c_out, int k, int stride, int padding, int act, float slope, float out_scale, float* y,
                         void* cuda_stream) {
  CWG_REQUIRE(batch >= 1 && c_in >= 1 && c_out >= 1 && k >= 1 && stride >= 1 && padding >= 0, "bad conv_transpose1d shape");
  const int t_out = (t_in - 1) * stride - 2 * padding + k;
  CWG_REQUIRE(t_in >= 1 && t_out >= 1, "conv_transpose1d: empty output");
  CWG_REQUIRE(act >= 0 && act <= 4, "conv_transpose1d: bad act");
  CWG_REQUIRE(x && w_phases && y, "NULL argument");
  const int jmax = (k + stride - 1) / stride;
  ConvP p{};
  p.B = batch; p.Cin = c_in; p.Tin = t_in; p.Cout = c_out; p.Tout = t_out; p.Kt = jmax; p.Q = t_in + jmax;
  p.tap_step = -1; p.in_off = 0; p.out_stride = stride; p.out_off = -padding; p.pad_mode = 0;
  p.act = act; p.slope = slope; p.out_scale = out_scale; p.x = x; p.w = w_phases; p.bias = bias; p.res = nullptr; p.y = y;
  p.w_phase_stride = (long long)c_out * c_in * jmax; p.out_off_phase_step = 1;
  const long long M = (long long)batch * p.Q;
  dim3 grid((unsigned)((M + CBM - 1) / CBM), (c_out + CBN - 1) / CBN, stride);
  k_cn_conv<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(p);
  CWG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int cwg_resample1d(const float* x, int batch, int channels, int t_in, long long x_batch_stride, float* y, int t_out,
                   long long y_batch_stride, int mode, int t_virtual, int crop, float scale_factor, int accumulate,
                   void* cuda_stream) {
  CWG_REQUIRE(batch >= 1 && channels >= 1 && t_in >= 1 && t_out >= 1 && t_virtual >= 1, "bad resample shape");
  CWG_REQUIRE(mode >= 0 && mode <= 2 && crop >= 0 && crop + t_out <= t_virtual, "resample: bad mode / crop");
  CWG_REQUIRE(x && y, "NULL argument");
  const long long n = (long long)batch * channels * t_out;
  k_cn_resample<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(
      x, batch, channels, t_in, x_batch_stride, y, t_out, y_batch_stride, mode, t_virtual, crop,
      scale_factor > 0.f ? 1.f / scale_factor : 0.f, accumulate);
  CWG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int cwg_deemphasis(const float* x, int batch, int t_samples, double coef, int vol_scaling, float* y, void* cuda_stream) {
  CWG_REQUIRE(batch >= 0 && t_samples >= 0, "negative size");
  if (batch == 0 || t_samples == 0) return 0;
  CWG_REQUIRE(x && y, "NULL argument");
  k_cn_deemph<<<batch, DE_THREADS, 0, (cudaStream_t)cuda_stream>>>(x, y, t_samples, coef, vol_scaling);
  CWG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"

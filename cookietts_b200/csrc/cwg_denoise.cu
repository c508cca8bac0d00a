// Denoiser post-filter on the GPU: the reference runs it on the CPU after every infer
// (denoiser.py:59-71 with the conv-based STFT of utils/audio/stft.py:79-146).
//   k_dn_gemm<0> : STFT analysis   frames(reflect-padded audio) x forward_basis^T -> [B*NF][2*cutoff] (re | im)
//   k_dn_scale   : magnitude - bias*strength, clamped at 0, re-applied to (re, im)   (the reference's
//                  atan2 / cos / sin round trip equals a per-bin scale of the complex value)
//   k_dn_gemm<1> : STFT synthesis  [B*NF][2*cutoff] x inverse_basis -> [B*NF][fl] frame contributions
//   k_dn_ola     : overlap-add of the frame contributions, window-sum-square normalisation,
//                  x filter_length/hop, trim filter_length/2 on both sides
// fp32 CUDA-core kernels: the filter costs ~0.15 % of the vocoder's MACs per sample.
#include "cwg_common.cuh"

namespace cwg {
namespace {

constexpr int DBM = 64, DBN = 64, DBK = 16;

struct DnGemmP {
  int M, N, K;
  const float* W;      // [N][K]
  const float* A;      // mode 0: audio [B][T]; mode 1: [M][K]
  float* C;            // [M][N]
  int T, NF, hop, pad; // mode 0
};

template <int MODE>
__device__ __forceinline__ float dn_load_a(const DnGemmP& p, int m, int k) {
  if (m >= p.M || k >= p.K) return 0.f;
  if (MODE == 0) {     // reflect padding of fl/2 on both sides (stft.py:92-96), frame f starts at f*hop
    int b = m / p.NF, f = m - b * p.NF;
    int i = f * p.hop + k - p.pad;
    if (i < 0) i = -i;
    if (i >= p.T) i = 2 * (p.T - 1) - i;
    return __ldg(p.A + (size_t)b * p.T + i);
  }
  return __ldg(p.A + (size_t)m * p.K + k);
}

template <int MODE>
__global__ void __launch_bounds__(256) k_dn_gemm(DnGemmP p) {
  __shared__ float As[DBK][DBM + 4];
  __shared__ float Bs[DBK][DBN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * DBM, n0 = blockIdx.y * DBN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += DBK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int idx = tid + r * 256;
      int row = idx / DBK, kk = idx % DBK;
      As[kk][row] = dn_load_a<MODE>(p, m0 + row, k0 + kk);
      int n = idx / DBK;
      Bs[kk][n] = (n0 + n < p.N && k0 + kk < p.K) ? __ldg(p.W + (size_t)(n0 + n) * p.K + k0 + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < DBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < p.M && n < p.N) p.C[(size_t)m * p.N + n] = acc[i][j];
    }
}

// in place on spec [M][2*cutoff]; bias [n_bias][cutoff], row chosen per utterance
__global__ void k_dn_scale(float* __restrict__ spec, const float* __restrict__ bias, const int* __restrict__ bias_index,
                           long long M, int NF, int cutoff, float strength) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * cutoff) return;
  long long m = i / cutoff; int j = (int)(i - m * cutoff);
  int b = (int)(m / NF);
  const float* brow = bias + (size_t)(bias_index ? bias_index[b] : 0) * cutoff;
  float* re = spec + m * 2 * cutoff + j; float* im = re + cutoff;
  const float r = *re, q = *im;
  const float mag = sqrtf(r * r + q * q);
  const float nw = fmaxf(mag - __ldg(brow + j) * strength, 0.f);     // denoiser.py:64-68
  if (mag > 0.f) { const float s = nw / mag; *re = r * s; *im = q * s; }
  else { *re = nw; *im = 0.f; }                                       // atan2(0, 0) = 0
}

// mean over frames of the magnitude, per utterance: out [B][cutoff]  (denoiser.py:50-57)
__global__ void k_dn_mean_mag(const float* __restrict__ spec, float* __restrict__ out, int NF, int cutoff) {
  int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cutoff) return;
  float acc = 0.f;
  for (int f = 0; f < NF; ++f) {
    const float* row = spec + ((size_t)b * NF + f) * 2 * cutoff;
    const float r = row[j], q = row[cutoff + j];
    acc += sqrtf(r * r + q * q);
  }
  out[(size_t)b * cutoff + j] = acc / (float)NF;
}

__global__ void k_dn_ola(const float* __restrict__ contrib, const float* __restrict__ window_sum, float* __restrict__ out,
                         int B, int NF, int fl, int hop, int T_out, float tiny, float scale) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T_out) return;
  int b = (int)(i / T_out), to = (int)(i - (long long)b * T_out);
  int t = to + fl / 2;                                                // stft.py:144-145 trim
  int f_hi = min(NF - 1, t / hop);
  int f_lo = max(0, (t - fl + hop) / hop);                            // smallest f with t - f*hop < fl
  float acc = 0.f;
  for (int f = f_lo; f <= f_hi; ++f) acc += contrib[((size_t)b * NF + f) * fl + (t - f * hop)];
  const float ws = __ldg(window_sum + t);
  if (ws > tiny) acc /= ws;                                           // stft.py:134-139
  out[i] = acc * scale;                                               // stft.py:142
}

struct DnDims { int B, T, fl, hop, cutoff, NF, T_out; long long M; };

int dn_dims(int batch, int n_samples, int fl, int hop, DnDims* d) {
  CWG_REQUIRE(batch >= 1 && fl >= 2 && hop >= 1 && n_samples > fl / 2, "bad denoiser shapes (need n_samples > filter_length/2)");
  d->B = batch; d->T = n_samples; d->fl = fl; d->hop = hop; d->cutoff = fl / 2 + 1;
  d->NF = (n_samples + 2 * (fl / 2) - fl) / hop + 1;
  d->T_out = fl + hop * (d->NF - 1) - 2 * (fl / 2);
  d->M = (long long)batch * d->NF;
  return 0;
}

int dn_analysis(const DnDims& d, const float* audio, const float* fwd_basis, float* spec, cudaStream_t s) {
  DnGemmP p{};
  p.M = (int)d.M; p.N = 2 * d.cutoff; p.K = d.fl; p.W = fwd_basis; p.A = audio; p.C = spec;
  p.T = d.T; p.NF = d.NF; p.hop = d.hop; p.pad = d.fl / 2;
  dim3 grid((p.M + DBM - 1) / DBM, (p.N + DBN - 1) / DBN);
  k_dn_gemm<0><<<grid, 256, 0, s>>>(p);
  CWG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace
}  // namespace cwg

using namespace cwg;

extern "C" {

size_t cwg_denoise_workspace_bytes(int batch, int n_samples, int filter_length, int hop_length) {
  DnDims d;
  if (dn_dims(batch, n_samples, filter_length, hop_length, &d)) return 0;
  return align_up((size_t)d.M * 2 * d.cutoff * 4, 1024) + align_up((size_t)d.M * d.fl * 4, 1024) + 1024;
}

int cwg_denoise_out_samples(int n_samples, int filter_length, int hop_length) {
  DnDims d;
  if (dn_dims(1, n_samples, filter_length, hop_length, &d)) return -1;
  return d.T_out;
}

int cwg_stft_mean_magnitude(const float* audio, int batch, int n_samples, int filter_length, int hop_length,
                            const float* fwd_basis, float* mean_mag, void* workspace, size_t workspace_bytes,
                            void* cuda_stream) {
  DnDims d;
  if (int r = dn_dims(batch, n_samples, filter_length, hop_length, &d)) return r;
  CWG_REQUIRE(audio && fwd_basis && mean_mag && workspace, "NULL argument");
  CWG_REQUIRE(workspace_bytes >= (size_t)d.M * 2 * d.cutoff * 4, "workspace too small");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  float* spec = (float*)workspace;
  if (int r = dn_analysis(d, audio, fwd_basis, spec, s)) return r;
  dim3 grid((d.cutoff + 127) / 128, d.B);
  k_dn_mean_mag<<<grid, 128, 0, s>>>(spec, mean_mag, d.NF, d.cutoff);
  CWG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int cwg_denoise(const float* audio, int batch, int n_samples, int filter_length, int hop_length,
                const float* fwd_basis, const float* inv_basis_t, const float* window_sum,
                const float* bias_spec, const int32_t* bias_index, float strength,
                float* out, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  DnDims d;
  if (int r = dn_dims(batch, n_samples, filter_length, hop_length, &d)) return r;
  CWG_REQUIRE(audio && fwd_basis && inv_basis_t && window_sum && bias_spec && out && workspace, "NULL argument");
  const size_t spec_bytes = align_up((size_t)d.M * 2 * d.cutoff * 4, 1024);
  CWG_REQUIRE(workspace_bytes >= spec_bytes + (size_t)d.M * d.fl * 4, "workspace too small");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  float* spec = (float*)workspace;
  float* contrib = (float*)((char*)workspace + spec_bytes);
  if (int r = dn_analysis(d, audio, fwd_basis, spec, s)) return r;
  {
    long long n = d.M * d.cutoff;
    k_dn_scale<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(spec, bias_spec, bias_index, d.M, d.NF, d.cutoff, strength);
    CWG_CHECK_CUDA(cudaGetLastError());
  }
  {
    DnGemmP p{};
    p.M = (int)d.M; p.N = d.fl; p.K = 2 * d.cutoff; p.W = inv_basis_t; p.A = spec; p.C = contrib;
    dim3 grid((p.M + DBM - 1) / DBM, (p.N + DBN - 1) / DBN);
    k_dn_gemm<1><<<grid, 256, 0, s>>>(p);
    CWG_CHECK_CUDA(cudaGetLastError());
  }
  {
    long long n = (long long)d.B * d.T_out;
    k_dn_ola<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(contrib, window_sum, out, d.B, d.NF, d.fl, d.hop, d.T_out,
                                                          1.17549435e-38f, (float)d.fl / (float)d.hop);
    CWG_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"

// CUDA-core (fp32 FFMA) kernels of the WaveGlow inverse pass, plus the flow-boundary kernel
// that both arithmetic modes share.  This path keeps exact fp32 semantics (CWG_MODE_FFMA) and
// doubles as the on-device cross-check for the tcgen05 path in cwg_tc.cu.
//
// Reference lines implemented (CookieTTS/_4_mtw/waveglow/glow.py):
//   cond GEMM        : upsample + squeeze + cond_layers[0..1]  :318-324, :198-199 (folded, see packing.py)
//   layer GEMM1+gate : in_layers[i] + cond_layers[2] slice + fused_add_tanh_sigmoid_multiply :201-209, :34-41
//   layer GEMM2      : res_skip_layers[i], residual/skip update, `end` (folded) :211-222
//   flow boundary    : affine coupling inverse, W^-1, early-z concat, start conv :329-347, :189
#include "cwg_common.cuh"

namespace cwg {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

struct GemmP {
  int M, N, K;
  const float* W; int ldw;          // [N][ldw]
  const float* a0; const float* a1; // A sources
  int Tm, Tp, nmel, C, ks, dil, H;
  float* o0; float* o1;
  const float* bias; const float* bias2; const float* xin;
  int bias_bstride;                 // cond: batch stride of the per-utterance bias
  int has_res, first;
};

// A(m, kk) for the three GEMMs.
template <int AMODE>
__device__ __forceinline__ float load_a(const GemmP& p, int m, int kk) {
  if (m >= p.M || kk >= p.K) return 0.f;
  if (AMODE == 0) {             // mel4[b, f, j*nmel + ci] = mel[b, ci, f - j]
    int b = m / p.Tm, f = m - b * p.Tm;
    int j = kk / p.nmel, ci = kk - j * p.nmel;
    int ff = f - j;
    return ff >= 0 ? __ldg(p.a0 + ((size_t)b * p.nmel + ci) * p.Tm + ff) : 0.f;
  } else if (AMODE == 1) {      // [x taps | h2]
    int kx = p.ks * p.C;
    if (kk < kx) {
      int b = m / p.Tp, t = m - b * p.Tp;
      int tap = kk / p.C, c = kk - tap * p.C;
      int tt = t + (tap - p.ks / 2) * p.dil;
      return (tt >= 0 && tt < p.Tp) ? __ldg(p.a0 + ((size_t)b * p.Tp + tt) * p.C + c) : 0.f;
    }
    return __ldg(p.a1 + (size_t)m * p.H + (kk - kx));
  } else {
    return __ldg(p.a0 + (size_t)m * p.K + kk);
  }
}

template <int EMODE>
__device__ __forceinline__ void store_c(const GemmP& p, int m, int n, float acc) {
  if (m >= p.M || n >= p.N) return;
  if (EMODE == 0) {             // H2[m][n] (+ per-utterance cond bias of channel n % H)
    int b = m / p.Tm;
    p.o0[(size_t)m * p.N + n] = acc + __ldg(p.bias + (size_t)b * p.bias_bstride + (n % p.H));
  } else if (EMODE == 1) {      // pre-activation
    p.o0[(size_t)m * p.N + n] = acc + __ldg(p.bias + n);
  } else {                      // res / folded end
    if (n < p.C) {
      if (p.has_res) p.o0[(size_t)m * p.C + n] = __ldg(p.xin + (size_t)m * p.C + n) + acc + __ldg(p.bias + n);
    } else {
      int j = n - p.C;
      float* e = p.o1 + (size_t)m * CWG_EO_PAD + j;
      *e = (p.first ? __ldg(p.bias2 + j) : *e) + acc;
    }
  }
}

template <int AMODE, int EMODE>
__global__ void __launch_bounds__(256) k_sgemm(GemmP p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  float acc[TM][TN] = {};
  for (int k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
    for (int r = 0; r < (BM * BK) / 256; ++r) {
      int idx = tid + r * 256;
      int row, kk;
      if (AMODE == 0) { kk = idx / BM; row = idx % BM; } else { row = idx / BK; kk = idx % BK; }
      As[kk][row] = load_a<AMODE>(p, m0 + row, k0 + kk);
    }
#pragma unroll
    for (int r = 0; r < (BN * BK) / 256; ++r) {
      int idx = tid + r * 256;
      int n = idx / BK, kk = idx % BK;
      Bs[kk][n] = (n0 + n < p.N && k0 + kk < p.K) ? __ldg(p.W + (size_t)(n0 + n) * p.ldw + k0 + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) store_c<EMODE>(p, m0 + ty * TM + i, n0 + tx * TN + j, acc[i][j]);
}

// acts[m][c] = tanh(pre[m][c]) * sigmoid(pre[m][C + c])   (glow.py:34-41)
__global__ void k_gate(const float* __restrict__ pre, float* __restrict__ acts, long long n, int C) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long m = i / C; int c = (int)(i - m * C);
  float a = pre[m * 2 * C + c], b = pre[m * 2 * C + C + c];
  acts[i] = tanhf(a) * (1.f / (1.f + expf(-b)));
}

struct BoundaryP {
  long long BT; int G, C;
  int init, do_flow, do_start;
  int n_rem, n_half;          // of flow_done
  int n_rem2, n_half2;        // of flow_next
  const float* z; float sigma;
  float* audio; const float* eo;
  const float* winv;          // [MAX_GROUP][MAX_GROUP] of flow_done
  const float* start_w;       // [C][MAX_GROUP/2] of flow_next
  const float* start_b;       // [C]
  void* x_out;
};

// One warp per group-step (a row of the [B*T'][G] audio state, which IS the [B, T] output buffer:
// latent channel c of step s lives at audio[b, s*G + c], so the early-z concat (glow.py:342-347)
// and the final un-squeeze (:349) are no-ops by layout).  Every lane redundantly evaluates the
// tiny coupling + W^-1 of its warp's row (broadcast loads), then the warp writes the next flow's
// `start` conv output: lane l owns channels [8l, 8l+8) of each 256-channel block, so a row of the
// bf16 planes is one coalesced 512-byte store per plane.
template <int XFMT>
__global__ void __launch_bounds__(256) k_flow_boundary(BoundaryP p) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  const int off2 = p.G - p.n_rem2;
  for (long long m = warp_global; m < p.BT; m += n_warps) {
    float a[CWG_MAX_GROUP];
#pragma unroll
    for (int g = 0; g < CWG_MAX_GROUP; ++g)
      a[g] = g < p.G ? (p.init ? p.sigma * __ldg(p.z + m * p.G + g) : p.audio[m * p.G + g]) : 0.f;
    if (p.do_flow) {
      const int off = p.G - p.n_rem;
      const float* e = p.eo + m * CWG_EO_PAD;
      float v[CWG_MAX_GROUP];
#pragma unroll
      for (int j = 0; j < CWG_MAX_GROUP; ++j) {
        float x = 0.f;
        if (j < p.n_rem) {
          x = a[off + j];
          if (j >= p.n_half) {     // audio_1 = (audio_1 - b) / exp(s), glow.py:337
            const float bb = __ldg(e + (j - p.n_half)), ss = __ldg(e + j);
            x = (x - bb) * expf(-ss);
          }
        }
        v[j] = x;
      }
#pragma unroll
      for (int r = 0; r < CWG_MAX_GROUP; ++r) {       // z = conv1d(z, W^-1), glow.py:98
        if (r < p.n_rem) {
          float acc = 0.f;
#pragma unroll
          for (int c = 0; c < CWG_MAX_GROUP; ++c)
            if (c < p.n_rem) acc = fmaf(__ldg(p.winv + r * CWG_MAX_GROUP + c), v[c], acc);
          a[off + r] = acc;
        }
      }
    }
    __syncwarp();               // every lane has read the row before any lane rewrites it
    if ((p.init || p.do_flow) && lane < p.G) {
      float mine = 0.f;
#pragma unroll
      for (int g = 0; g < CWG_MAX_GROUP; ++g) if (g == lane) mine = a[g];
      p.audio[m * p.G + lane] = mine;
    }
    if (!p.do_start) continue;
    float a0[CWG_MAX_GROUP / 2];
#pragma unroll
    for (int j = 0; j < CWG_MAX_GROUP / 2; ++j) {
      float x = 0.f;
#pragma unroll
      for (int g = 0; g < CWG_MAX_GROUP; ++g) if (g == off2 + j && j < p.n_half2) x = a[g];
      a0[j] = x;
    }
    for (int c0 = lane * 8; c0 < p.C; c0 += 256) {      // audio = start(audio_0), glow.py:189
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = c0 + i;
        float acc = 0.f;
        if (c < p.C) {
          acc = __ldg(p.start_b + c);
          const float* w = p.start_w + (size_t)c * (CWG_MAX_GROUP / 2);
#pragma unroll
          for (int j = 0; j < CWG_MAX_GROUP / 2; ++j)
            if (j < p.n_half2) acc = fmaf(__ldg(w + j), a0[j], acc);
        }
        x[i] = acc;
      }
      const size_t idx = (size_t)m * p.C + c0;
      if (XFMT == 0) {
        float* o = reinterpret_cast<float*>(p.x_out) + idx;
        if (c0 + 8 <= p.C && (p.C & 3) == 0) {
          *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(x[4], x[5], x[6], x[7]);
        } else {
          for (int i = 0; i < 8; ++i) if (c0 + i < p.C) o[i] = x[i];
        }
      } else {
        __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(p.x_out) + idx;
        __nv_bfloat16* lo = hi + (size_t)p.BT * p.C;
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
          __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * i] - __low2float(hh), x[2 * i + 1] - __high2float(hh));
          h[i] = *reinterpret_cast<uint32_t*>(&hh);
          l[i] = *reinterpret_cast<uint32_t*>(&ll);
        }
        if (c0 + 8 <= p.C && (p.C & 7) == 0) {
          *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
        } else {
          for (int i = 0; i < 8; ++i)
            if (c0 + i < p.C) {
              __nv_bfloat16 hv = __float2bfloat16_rn(x[i]);
              hi[i] = hv;
              lo[i] = __float2bfloat16_rn(x[i] - __bfloat162float(hv));
            }
        }
      }
    }
  }
}

}  // namespace

int launch_cond_ffma(const Dims& d, const cwg_weights* w, int flow, const float* mel,
                     const float* cond_bias, float* h2, cudaStream_t s) {
  GemmP p{};
  p.M = d.B * d.Tm; p.N = d.P * d.H; p.K = d.KC;
  p.W = w->cond_w_f32 + (size_t)flow * p.N * p.K; p.ldw = p.K;
  p.a0 = mel; p.Tm = d.Tm; p.nmel = d.M; p.H = d.H;
  p.o0 = h2; p.bias = cond_bias + (size_t)flow * d.H; p.bias_bstride = d.F * d.H;
  dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN);
  k_sgemm<0, 0><<<grid, 256, 0, s>>>(p);
  CWG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_layer_ffma(const Dims& d, const cwg_weights* w, int flow, int layer, const float* x_in,
                      float* x_out, const float* h2, float* eo, float* pre, float* acts, cudaStream_t s) {
  const size_t fl = (size_t)flow * d.L + layer;
  GemmP p{};
  p.M = (int)d.BT; p.N = 2 * d.C; p.K = d.K1;
  p.W = w->w1_f32 + fl * (size_t)(2 * d.C) * d.K1; p.ldw = d.K1;
  p.a0 = x_in; p.a1 = h2; p.Tp = d.Tp; p.C = d.C; p.ks = d.ks; p.dil = 1 << layer; p.H = d.H;
  p.o0 = pre; p.bias = w->b1 + fl * (size_t)(2 * d.C);
  dim3 g1((p.M + BM - 1) / BM, (p.N + BN - 1) / BN);
  k_sgemm<1, 1><<<g1, 256, 0, s>>>(p);
  CWG_CHECK_CUDA(cudaGetLastError());

  long long n = d.BT * d.C;
  k_gate<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pre, acts, n, d.C);
  CWG_CHECK_CUDA(cudaGetLastError());

  GemmP q{};
  q.M = (int)d.BT; q.N = d.N2; q.K = d.C;
  q.W = w->w2_f32 + fl * (size_t)d.N2 * d.C; q.ldw = d.C;
  q.a0 = acts; q.C = d.C;
  q.o0 = x_out; q.o1 = eo; q.xin = x_in;
  q.bias = w->b2 + fl * (size_t)d.C; q.bias2 = w->eo_b + (size_t)flow * CWG_EO_PAD;
  q.has_res = layer < d.L - 1; q.first = layer == 0;
  dim3 g2((q.M + BM - 1) / BM, (q.N + BN - 1) / BN);
  k_sgemm<2, 2><<<g2, 256, 0, s>>>(q);
  CWG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_flow_boundary(const cwg_config* cfg, const Dims& d, const cwg_weights* w, int xfmt,
                         int flow_done, int flow_next, const float* z, float sigma, float* audio,
                         const float* eo, void* x_out, cudaStream_t s) {
  BoundaryP p{};
  p.BT = d.BT; p.G = d.G; p.C = d.C;
  p.init = z != nullptr; p.do_flow = flow_done >= 0; p.do_start = flow_next >= 0;
  p.z = z; p.sigma = sigma; p.audio = audio; p.eo = eo; p.x_out = x_out;
  if (p.do_flow) {
    flow_channels(cfg, flow_done, &p.n_rem, &p.n_half);
    p.winv = w->winv + (size_t)flow_done * CWG_MAX_GROUP * CWG_MAX_GROUP;
  }
  if (p.do_start) {
    flow_channels(cfg, flow_next, &p.n_rem2, &p.n_half2);
    p.start_w = w->start_w + (size_t)flow_next * d.C * (CWG_MAX_GROUP / 2);
    p.start_b = w->start_b + (size_t)flow_next * d.C;
  }
  // 8 warps per block, one group-step per warp per iteration; enough blocks for ~16 rows per warp
  long long want = (d.BT + 8 * 16 - 1) / (8 * 16);
  unsigned grid = (unsigned)(want < 148 ? 148 : (want > 148 * 64 ? 148 * 64 : want));
  if (xfmt == 0) k_flow_boundary<0><<<grid, 256, 0, s>>>(p);
  else           k_flow_boundary<1><<<grid, 256, 0, s>>>(p);
  CWG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cwg

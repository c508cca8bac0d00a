// WaveFlow inverse pass on the fp32 CUDA cores (CWG_MODE_FFMA of cwg_wf_infer): the exact-fp32 cross-check of the
// tensor-core WaveFlow kernels (cwg_wf.cu) and the path for every WN_2d shape those are not specialised for - any even
// channel count, any kernel (kernel_h x odd kernel_w, e.g. the 7x7 depthwise-separable in_layers of the reference's trained
// WaveFlow checkpoints, folded to a dense conv at pack time), squeeze heights up to 32, any number of cond channels.
// Same algorithm, same packed-weight layout and same host loop as cwg_wf.cu:
//   efficient_model_ax.py:279-357 WaveGlow.inverse, efficient_modules.py:42-65 WaveFlowCoupling.inverse (row by row),
//   glow_ax.py:556-635 WN_2d.forward with its per-layer conv queue (a ring of kernel_h rows per layer here),
//   efficient_modules.py:360-403 PermuteHeight (host-side column bookkeeping).
// WN_config variants (ABI 5): every gated unit of glow_ax.py:168-198, listed width / height dilations (:513-517; the ring
// of a layer holds (kernel_h - 1) * dilation_h + 1 rows), merge_res_skip / res_skip=False (packing only: zero res rows),
// WN-level speaker embeddings (a per-utterance gate bias, cwg_wf_weights.b1_batch), early outputs (efficient_model_ax.py:
// 319-322,:339-340: a flow works on the trailing n_rem height rows only) and mix_first = 0 (PermuteHeight before the coupling).
#include "cwg_common.cuh"

namespace cwg {
namespace {

constexpr int GM = 128, GN = 64, GK = 16;
constexpr int WFF_MAX_GROUP = 32;

struct WffGemmP {
  long long BT; int Tp, C, KH, KW, M;      // M = cond channels
  int dil_h, ring_rows; long long bias_bstride;   // ring_rows = (KH - 1) * max dil_h + 1; bias_bstride: per-utterance b1 (0: shared)
  const float* c_add; long long c_bstride;        // GEMM1: caller-evaluated cond slice [B][..][Tp] added to the pre-activation, or NULL
  int N, K;                                // output columns, contraction length
  const float* W; const float* bias;       // W [N][K]
  // GEMM1 (in_layer + cond): x ring of this layer [KH][BT][C]; newest row index `row`; mel_up [BT][M]
  const float* ring; const float* mel; int row, dil;
  float* pre;                              // [BT][2C]
  // GEMM2 (res_skip + folded end): acts [BT][C]; x_cur = ring slot of `row`; x_next = next layer's slot (NULL: last layer)
  const float* acts; const float* x_cur; float* x_next; float* eo; const float* eo_b; int first;
};

// Eight consecutive k (k1 .. k1+7) of row m of the implicit A matrix.  MODE 0 (GEMM1): k = ((a * KW + b) * C + c) over the
// conv-queue rows a (causal in height, zero queue: glow_ax.py:597-602) and width taps b ('same' zero padding), then the M cond
// columns; channels are contiguous in memory, so the (a, b) decomposition is redone only when a run crosses a tap.
// MODE 1 (GEMM2): k = channel of acts.
template <int MODE>
__device__ __forceinline__ void wff_a8(const WffGemmP& p, long long m, int k1, float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (m >= p.BT) return;
  if (MODE == 1) {
    const float* src = p.acts + (size_t)m * p.C;
#pragma unroll
    for (int i = 0; i < 8; ++i) if (k1 + i < p.K) v[i] = __ldg(src + k1 + i);
    return;
  }
  const int kx = p.KH * p.KW * p.C;
  const long long ub = m / p.Tp; const int t = (int)(m - ub * p.Tp);
  int seg = -1, c = 0; const float* src = nullptr;          // current (a, b) tap: base pointer of its channel run, or NULL = zeros
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int kk = k1 + i;
    if (kk >= p.K) break;
    if (kk >= kx) { v[i] = __ldg(p.mel + (size_t)m * p.M + (kk - kx)); continue; }
    if (seg < 0 || c >= p.C) {
      seg = kk / p.C; c = kk - seg * p.C;
      const int a = seg / p.KW, b = seg - a * p.KW;
      const int src_row = p.row - (p.KH - 1 - a) * p.dil_h;
      const int tt = t + (b - p.KW / 2) * p.dil;
      src = (src_row < 0 || tt < 0 || tt >= p.Tp) ? nullptr
            : p.ring + ((size_t)(src_row % p.ring_rows) * p.BT + (size_t)ub * p.Tp + tt) * p.C;
    }
    if (src) v[i] = __ldg(src + c);
    ++c;
  }
}

// 128 x 64 block tile, 256 threads, 8 rows (m = tx*4 + i and 64 + tx*4 + i) x 4 columns per thread, LDS.128 operand reads,
// the next k-tile fetched into registers while the current one is multiplied.
template <int MODE>     // 0: GEMM1, 1: GEMM2
__global__ void __launch_bounds__(256, 2) k_wff_gemm(WffGemmP p) {
  __shared__ __align__(16) float As[GK][GM + 4];
  __shared__ __align__(16) float Bs[GK][GN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const long long m0 = (long long)blockIdx.x * GM;
  const int n0 = blockIdx.y * GN;
  float acc[8][4] = {};
  const int a_row = tid >> 1, a_k = (tid & 1) * 8;           // A: row a_row, k a_k .. a_k+7 of the tile
  const int b_n = tid >> 2, b_k = (tid & 3) * 4;             // B: row b_n, k b_k .. b_k+3
  float ra[8], rb[4];
  auto fetch = [&](int k0) {
    wff_a8<MODE>(p, m0 + a_row, k0 + a_k, ra);
    const float* wr = p.W + (size_t)(n0 + b_n) * p.K + k0 + b_k;
#pragma unroll
    for (int i = 0; i < 4; ++i) rb[i] = (n0 + b_n < p.N && k0 + b_k + i < p.K) ? __ldg(wr + i) : 0.f;
  };
  fetch(0);
  for (int k0 = 0; k0 < p.K; k0 += GK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[a_k + i][a_row] = ra[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[b_k + i][b_n] = rb[i];
    __syncthreads();
    if (k0 + GK < p.K) fetch(k0 + GK);
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tx * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + tx * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
    if (m >= p.BT) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + ty * 4 + j;
      if (n >= p.N) continue;
      if (MODE == 0) {
        const long long ub = m / p.Tp;
        float v = acc[i][j] + __ldg(p.bias + ub * p.bias_bstride + n);
        if (p.c_add) v += __ldg(p.c_add + ub * p.c_bstride + (size_t)n * p.Tp + (m - ub * p.Tp));
        p.pre[(size_t)m * p.N + n] = v;
      } else if (n < p.C) {                                       // x = x + res (glow_ax.py:620-626)
        if (p.x_next) p.x_next[(size_t)m * p.C + n] = __ldg(p.x_cur + (size_t)m * p.C + n) + acc[i][j] + __ldg(p.bias + n);
      } else {                                                    // folded `end` of the skip path: (log_s, t)
        const int e = n - p.C;
        float* q = p.eo + (size_t)m * CWG_EO_PAD + e;
        *q = (p.first ? __ldg(p.eo_b + e) : *q) + acc[i][j];
      }
    }
  }
}

// Depthwise part of a separable in_layer (glow_ax.py:525-528): d[m][c] = b[c] + sum_{a,b} w[c][a][b] * x_queue[row - (KH-1-a)*dil_h]
// [t + (b - KW/2)*dil_w][c], causal in height (zero queue), 'same' zero padding in width.  One thread per (group-step, channel).
__global__ void k_wff_depthwise(long long BT, int Tp, int C, int KH, int KW, const float* __restrict__ ring, int row, int dil_w,
                                int dil_h, int ring_rows, const float* __restrict__ w, const float* __restrict__ bias,
                                float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BT * C) return;
  const long long m = i / C; const int c = (int)(i - m * C);
  const long long ub = m / Tp; const int t = (int)(m - ub * Tp);
  float acc = __ldg(bias + c);
  const float* wc = w + (size_t)c * KH * KW;
  for (int a = 0; a < KH; ++a) {
    const int src_row = row - (KH - 1 - a) * dil_h;
    if (src_row < 0) continue;
    const float* base = ring + ((size_t)(src_row % ring_rows) * BT + (size_t)ub * Tp) * C + c;
    for (int b = 0; b < KW; ++b) {
      const int tt = t + (b - KW / 2) * dil_w;
      if (tt >= 0 && tt < Tp) acc = fmaf(__ldg(wc + a * KW + b), __ldg(base + (size_t)tt * C), acc);
    }
  }
  out[i] = acc;
}

__global__ void k_wff_gate(const float* __restrict__ pre, float* __restrict__ acts, long long n, int C, int gate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long m = i / C; const int c = (int)(i - m * C);
  const float a = pre[m * 2 * C + c], b = pre[m * 2 * C + C + c];
  acts[i] = gate == CWG_GATE_GTU ? tanhf(a) * (1.f / (1.f + expf(-b))) : gated_unit(gate, a, b);   // glow_ax.py:36-166
}

__global__ void k_wff_mel_up(const float* __restrict__ mel, float* __restrict__ out, int B, int M, int frames,
                             int frames_padded, int Tp, int linear) {
  const long long n = (long long)B * Tp * M;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % M); const long long m = i / M;
  const int t = (int)(m % Tp), b = (int)(m / Tp);
  const float* row = mel + ((size_t)b * M + c) * frames;
  auto at = [&](int f) { return f < frames ? row[f] : 0.f; };
  float v;
  if (linear) {                                                   // F.interpolate(mode='linear', align_corners=True)
    const double src = Tp > 1 ? (double)t * (double)(frames_padded - 1) / (double)(Tp - 1) : 0.0;
    const int i0 = min((int)floor(src), frames_padded - 1), i1 = min(i0 + 1, frames_padded - 1);
    const float w = (float)(src - (double)i0);
    v = at(i0) * (1.f - w) + at(i1) * w;
  } else {
    v = at(min((int)floor((double)t * ((double)frames_padded / (double)Tp)), frames_padded - 1));
  }
  out[i] = v;
}

// finishes AR step row-1 and prepares step row (see k_wf_row in cwg_wf.cu)
__global__ void k_wff_row(long long BT, int G, int C, const float* __restrict__ in, float in_scale, int col_in,
                          float* __restrict__ out, int col_out, const float* __restrict__ eo, int do_couple,
                          const float* __restrict__ start_w, const float* __restrict__ start_b, float* __restrict__ x0) {
  const long long m = blockIdx.x;
  __shared__ float vs;
  if (threadIdx.x == 0) {
    float v = in[m * G + col_in] * in_scale;
    if (do_couple) v = (v - eo[m * CWG_EO_PAD + 1]) * expf(-eo[m * CWG_EO_PAD]);   // efficient_modules.py:62-63
    out[m * G + col_out] = v;
    vs = v;
  }
  if (x0 == nullptr) return;
  __syncthreads();
  const float v = vs;
  for (int c = threadIdx.x; c < C; c += blockDim.x) x0[(size_t)m * C + c] = fmaf(__ldg(start_w + c), v, __ldg(start_b + c));
}

__global__ void k_wff_scale(const float* __restrict__ in, float* __restrict__ out, long long n, float scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * scale;
}

// InvertibleConv1x1.inverse over the active height rows (efficient_modules.py:269-286): out[off + c] = sum_j Winv[c][j] in[off + j]
__global__ void k_wff_mix(const float* __restrict__ in, float* __restrict__ out, long long BT, int G, int off, int h,
                          const float* __restrict__ winv) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= BT) return;
  float v[WFF_MAX_GROUP];
  for (int j = 0; j < h; ++j) v[j] = in[m * G + off + j];
  for (int c = 0; c < h; ++c) {
    float acc = 0.f;
    for (int j = 0; j < h; ++j) acc = fmaf(__ldg(winv + c * WFF_MAX_GROUP + j), v[j], acc);
    out[m * G + off + c] = acc;
  }
  if (out != in) for (int j = 0; j < off; ++j) out[m * G + j] = in[m * G + j];      // early rows ride along to the waveform
}

// rows the flow k works on (efficient_model_ax.py:151-167)
inline int wff_n_rem(const cwg_wf_config* c, int k) {
  int n = c->n_group;
  if (c->n_early_every > 0)
    for (int j = 1; j <= k; ++j) if (j % c->n_early_every == 0) n -= c->n_early_size;
  return n;
}

void wff_perm(int k, int h, int* idx) {               // PermuteHeight index list of flow k (efficient_modules.py:341-353)
  if (k % 4 == 2 || k % 4 == 3) {
    const int half = h / 2;
    for (int i = 0; i < half; ++i) idx[i] = half - 1 - i;
    for (int i = half; i < h; ++i) idx[i] = h - 1 - (i - half);
  } else {
    for (int i = 0; i < h; ++i) idx[i] = h - 1 - i;
  }
}

inline int wff_dil_w(const cwg_wf_config* c, int l) { return c->dilations_w[l] > 0 ? c->dilations_w[l] : 1 << l; }
inline int wff_dil_h(const cwg_wf_config* c, int l) { return c->dilations_h[l] > 0 ? c->dilations_h[l] : 1; }
inline int wff_ring_rows(const cwg_wf_config* c) {
  int mx = 1;
  for (int l = 0; l < c->n_layers; ++l) mx = wff_dil_h(c, l) > mx ? wff_dil_h(c, l) : mx;
  return (c->kernel_h - 1) * mx + 1;
}

struct WffWs { float *eo, *state, *mel_up, *x, *pre, *acts, *dwo; size_t bytes; };
void wff_carve(const cwg_wf_config* c, long long BT, void* base, WffWs* ws) {
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n * sizeof(float), 1024); return (float*)((char*)base + o); };
  ws->eo = take((size_t)BT * CWG_EO_PAD); ws->state = take((size_t)BT * c->n_group); ws->mel_up = take((size_t)BT * c->n_mel);
  ws->x = take((size_t)c->n_layers * wff_ring_rows(c) * BT * c->n_channels);
  ws->pre = take((size_t)BT * 2 * c->n_channels); ws->acts = take((size_t)BT * c->n_channels);
  ws->dwo = take((size_t)BT * c->n_channels);                      // depthwise output of a separable in_layer
  ws->bytes = off;
}

}  // namespace

int wff_check(const cwg_wf_config* c, int batch, int t_samples) {
  CWG_REQUIRE(c != nullptr, "cfg is NULL");
  CWG_REQUIRE(c->n_channels >= 2 && c->n_channels % 2 == 0 && c->kernel_h >= 1 && c->kernel_h <= 16 && c->kernel_w >= 1 &&
              c->kernel_w % 2 == 1 && c->kernel_w <= 15, "fp32 WaveFlow path: n_channels even, 1 <= kernel_h <= 16, odd kernel_w <= 15");
  CWG_REQUIRE(c->n_group >= 2 && c->n_group <= WFF_MAX_GROUP, "n_group must be in [2, %d]", WFF_MAX_GROUP);
  CWG_REQUIRE(c->n_mel >= 1 && c->n_flows >= 1 && c->n_layers >= 1 && c->n_layers <= 16, "bad n_mel / n_flows / n_layers");
  CWG_REQUIRE(batch >= 1 && t_samples >= c->n_group && t_samples % c->n_group == 0, "t_samples must be a positive multiple of n_group");
  CWG_REQUIRE(c->gate >= 0 && c->gate < CWG_GATE_COUNT, "unknown gated unit %d", c->gate);
  for (int l = 0; l < c->n_layers; ++l)
    CWG_REQUIRE(c->dilations_w[l] >= 0 && c->dilations_h[l] >= 0 && c->dilations_h[l] <= 64, "bad dilations of layer %d", l);
  CWG_REQUIRE(c->n_early_every >= 0 && c->n_early_size >= 0 && wff_n_rem(c, c->n_flows - 1) >= 2, "too many early outputs for n_group");
  return 0;
}

size_t wff_workspace_bytes(const cwg_wf_config* c, int batch, int t_samples) {
  if (wff_check(c, batch, t_samples)) return 0;
  WffWs ws;
  wff_carve(c, (long long)batch * (t_samples / c->n_group), nullptr, &ws);
  return ws.bytes + 1024;
}

int wff_launch_count(const cwg_wf_config* c) {
  int n = 1 + (wff_n_rem(c, c->n_flows - 1) < c->n_group ? 1 : 0) + (c->mixing_conv ? c->n_flows + 1 : 0);
  for (int k = 0; k < c->n_flows; ++k) { const int h = wff_n_rem(c, k); n += h + (h - 1) * c->n_layers * 3; }   // (+1 per layer when separable)
  return n;
}

int wff_infer(const cwg_wf_config* cfg, const cwg_wf_weights* w, const float* mel, int frames, int pad_frames,
              const float* z, float sigma, float* audio, void* workspace, size_t workspace_bytes, int batch, int t_samples,
              cudaStream_t s, void** ev_begin, void** ev_end, int n_events) {
  if (int r = wff_check(cfg, batch, t_samples)) return r;
  CWG_REQUIRE(w && w->w1_f32 && w->w2_f32 && w->b1 && w->b2 && w->eo_b && w->start_w && w->start_b, "fp32 weight arrays missing");
  const int h = cfg->n_group, F = cfg->n_flows, L = cfg->n_layers, C = cfg->n_channels, KH = cfg->kernel_h, KW = cfg->kernel_w;
  const int M = w->c_all ? 0 : cfg->n_mel;                         // caller-evaluated cond path: no cond columns in w1
  const int Tp = t_samples / h;
  const long long BT = (long long)batch * Tp;
  WffWs ws;
  wff_carve(cfg, BT, workspace, &ws);
  CWG_REQUIRE(ws.bytes <= workspace_bytes, "workspace too small: need %zu, got %zu", ws.bytes, workspace_bytes);
  const bool sep = w->dw_w != nullptr;                             // separable in_layers: depthwise kernel, then a 1x1 GEMM
  CWG_REQUIRE(!sep || w->dw_b, "dw_b missing");
  const int K1 = (sep ? C : KH * KW * C) + M, N2 = C + CWG_EO_PAD;
  const int R = wff_ring_rows(cfg);                                // rows of every layer's conv queue
  if (M > 0) {
    const long long n = BT * M;
    k_wff_mel_up<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mel, ws.mel_up, batch, M, frames, frames + pad_frames, Tp, cfg->upsample_linear);
    CWG_CHECK_CUDA(cudaGetLastError());
  }
  // phys[c]: column of the [BT][G] state that holds logical height row c.  A flow with early outputs in front of it
  // (efficient_model_ax.py:319-322,:339-340) works on the trailing n_rem logical rows only.
  int phys[WFF_MAX_GROUP], perm[WFF_MAX_GROUP], nxt[WFF_MAX_GROUP];
  const int G = cfg->n_group;
  for (int c = 0; c < G; ++c) phys[c] = c;
  const size_t slot = (size_t)BT * C;                              // one ring row of one layer
  const bool conv_mix = cfg->mixing_conv != 0;                     // dense mixing: rows stay where they are (phys = identity)
  CWG_REQUIRE(!conv_mix || w->winv, "mixing_conv needs cwg_wf_weights.winv");
  const bool early = wff_n_rem(cfg, F - 1) < G || conv_mix;        // (the dense mixing reads / writes the state buffer)
  if (early) {                                                     // rows the first flow does not touch must be sigma * z too
    const long long n = BT * G;
    k_wff_scale<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(z, ws.state, n, sigma);
    CWG_CHECK_CUDA(cudaGetLastError());
  }
  auto permute = [&](int h, int off) {
    for (int c = 0; c < h; ++c) nxt[c] = phys[off + perm[c]];
    for (int c = 0; c < h; ++c) phys[off + c] = nxt[c];
  };
  auto mix = [&](int k, int h, int off, float* dst) -> int {
    k_wff_mix<<<(unsigned)((BT + 127) / 128), 128, 0, s>>>(ws.state, dst, BT, G, off, h,
                                                          w->winv + (size_t)k * WFF_MAX_GROUP * WFF_MAX_GROUP);
    CWG_CHECK_CUDA(cudaGetLastError());
    return 0;
  };
  int ev = 0;
  for (int k = F - 1; k >= 0; --k) {
    const int h = wff_n_rem(cfg, k), off = G - h;                  // active rows: logical off .. G-1
    wff_perm(k, h, perm);
    const bool from_z = k == F - 1 && !early, last_flow = k == 0;
    // the last flow writes the waveform itself - unless a dense mixing still follows its coupling
    const bool to_audio = last_flow && !(conv_mix && !cfg->mix_first_off);
    if (cfg->mix_first_off) {                                      // mix_first = 0: mixing inverse before the coupling
      if (conv_mix) { if (int r = mix(k, h, off, ws.state)) return r; } else permute(h, off);
    }
    for (int i = -1; i < h - 1; ++i) {
      if (i >= 0) {
        for (int l = 0; l < L; ++l, ++ev) {
          if (ev < n_events) CWG_CHECK_CUDA(cudaEventRecord((cudaEvent_t)ev_begin[ev], s));
          const size_t idx = (size_t)k * L + l;
          WffGemmP p{};
          p.BT = BT; p.Tp = Tp; p.C = C; p.KH = KH; p.KW = KW; p.M = M;
          p.N = 2 * C; p.K = K1; p.W = w->w1_f32 + idx * 2 * C * K1; p.bias = w->b1 + idx * 2 * C;
          p.ring = ws.x + (size_t)l * R * slot; p.mel = ws.mel_up; p.row = i; p.dil = wff_dil_w(cfg, l); p.pre = ws.pre;
          p.dil_h = wff_dil_h(cfg, l); p.ring_rows = R;
          if (sep) {
            const long long n = BT * C;
            k_wff_depthwise<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(BT, Tp, C, KH, KW, p.ring, i, p.dil, p.dil_h, R,
                                                                          w->dw_w + idx * C * KH * KW, w->dw_b + idx * C, ws.dwo);
            p.KH = 1; p.KW = 1; p.ring = ws.dwo; p.row = 0; p.dil = 1; p.dil_h = 1; p.ring_rows = 1;   // pointwise 1x1 on d
          }
          if (w->b1_batch) { p.bias = w->b1_batch + idx * 2 * C; p.bias_bstride = (long long)F * L * 2 * C; }
          if (w->c_all) {
            p.c_bstride = (long long)2 * C * L * Tp;
            p.c_add = w->c_all + (size_t)k * batch * p.c_bstride + (size_t)2 * C * l * Tp;
          }
          dim3 g1((unsigned)((BT + GM - 1) / GM), (unsigned)((2 * C + GN - 1) / GN));
          k_wff_gemm<0><<<g1, 256, 0, s>>>(p);
          const long long n = BT * C;
          k_wff_gate<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ws.pre, ws.acts, n, C, cfg->gate);
          WffGemmP q{};
          q.BT = BT; q.Tp = Tp; q.C = C; q.N = N2; q.K = C; q.W = w->w2_f32 + idx * N2 * C; q.bias = w->b2 + idx * C;
          q.acts = ws.acts; q.x_cur = ws.x + ((size_t)l * R + (i % R)) * slot;
          q.x_next = l < L - 1 ? ws.x + ((size_t)(l + 1) * R + (i % R)) * slot : nullptr;
          q.eo = ws.eo; q.eo_b = w->eo_b + (size_t)k * CWG_EO_PAD; q.first = l == 0;
          dim3 g2((unsigned)((BT + GM - 1) / GM), (unsigned)((C + 2 + GN - 1) / GN));     // res rows + the 2 folded-end rows
          q.N = C + 2;
          k_wff_gemm<1><<<g2, 256, 0, s>>>(q);
          CWG_CHECK_CUDA(cudaGetLastError());
          if (ev < n_events) CWG_CHECK_CUDA(cudaEventRecord((cudaEvent_t)ev_end[ev], s));
        }
      }
      const int j = i + 1;                                         // active row produced now (logical row off + j)
      float* x0 = j < h - 1 ? ws.x + (size_t)(j % R) * slot : nullptr;         // ring of layer 0
      // the last flow (h = G, off = 0) writes the waveform in its final logical order: after the coupling the mixing moves
      // row j to perm[j] (an involution); with mix_first = 0 nothing follows the coupling
      const int col_out = to_audio ? ((cfg->mix_first_off || conv_mix) ? j : perm[j]) : phys[off + j];
      k_wff_row<<<(unsigned)BT, 128, 0, s>>>(BT, G, C, from_z ? z : ws.state, from_z ? sigma : 1.f, phys[off + j],
                                            to_audio ? audio : ws.state, col_out, ws.eo, i >= 0,
                                            w->start_w + (size_t)k * C, w->start_b + (size_t)k * C, x0);
      CWG_CHECK_CUDA(cudaGetLastError());
    }
    if (!cfg->mix_first_off) {                                     // mix_first: mixing inverse after the coupling
      if (conv_mix) { if (int r = mix(k, h, off, last_flow ? audio : ws.state)) return r; } else permute(h, off);
    }
  }
  return 0;
}

}  // namespace cwg

// Mel-domain flow decoder, inverse pass: the reference's FlowDecoder.inverse
//   CookieTTS/_2_ttm/flowtts/waveglow/glow.py:302-343 (flows in reverse, mix_first, early z), WN.forward :133-172,
//   AffineCouplingBlock.inverse modules.py:36-47, InvertibleConv1x1.inverse modules.py:234-250,
//   and the untts variant (untts/waveglow/glow.py:126-127: constant padding of the first flow's hidden tensor).
// These decoders run at mel-frame rate (a few hundred steps per utterance, n_group up to 256 channels), three to four
// orders of magnitude less work than the vocoder, so they are built as fp32 CUDA-core kernels on the reference's own
// channels-first [B, C, T] layout: one implicit-GEMM conv kernel with four epilogues (store, accumulate, GTU gate, and
// the W^-1 mixing as a 1x1 conv) plus the coupling update.  Exact fp32 semantics; no tensor-core mode is needed here.
//
// The same kernels serve the general fp32 mode of the ax vocoder's 1-D WN (cwg_axg_flow, include/cwg.h): the WN_config
// variants of glow_ax.py:245-418 outside the packed layer kernels' specialisation - the 14 gated units of glow_ax.py:36-198,
// listed dilations, merge_res_skip / res_skip=False, multi-layer cond stacks (evaluated by the caller).
#include "cwg_common.cuh"

namespace cwg {
namespace {

constexpr int FBM = 128, FBN = 64, FBK = 16;

struct FdConvP {
  int B, Cin, T, N, ks, dil;
  float pad_value;
  int gate;                                   // EPI 2: CWG_GATE_* (0 = GTU)
  const float* x; long long x_bstride;        // x[b][ci][t] = x[b * x_bstride + ci * T + t]
  const float* w; const float* bias;          // w [rows][Cin][ks]; EPI 2 also reads rows n + N
  const float* add; long long add_bstride;    // EPI 2: cond slice [b][2N][t] added to the pre-activation
  float* y; long long y_bstride;              // y[b][n][t]
};

// Implicit-GEMM conv1d on [B, C, T]: rows m = (b, t), columns n = output channel, K ordered TAP-MAJOR (k = tap * Cin + ci) so
// that a 16-deep k-tile normally lies inside one tap and the gather needs no per-element division.  128 x 64 block tile,
// 256 threads, 8 (rows, as two groups of 4: m = tx*4 + i and 64 + tx*4 + i, conflict-free LDS.128) x 4 (columns) per thread;
// the next k-tile is fetched into registers while the current one is multiplied.
// EPI 0: y = conv + bias;  1: y += conv + bias;  2: y = gated_unit(pre[n], pre[n + N]), pre = conv + bias + add
template <int EPI>
__global__ void __launch_bounds__(256, 2) k_fd_conv(FdConvP p) {
  __shared__ __align__(16) float As[FBK][FBM + 4];
  __shared__ __align__(16) float Bs[FBK][FBN + 4];
  __shared__ __align__(16) float Bg[EPI == 2 ? FBK : 1][FBN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const long long M = (long long)p.B * p.T;
  const long long m0 = (long long)blockIdx.x * FBM;
  const int n0 = blockIdx.y * FBN;
  const int KD = p.Cin * p.ks;
  float acc[8][4] = {}, acg[EPI == 2 ? 8 : 1][4] = {};
  const int a_row = tid % FBM, a_k = tid / FBM;              // a_k in {0, 1}: this thread gathers k = a_k + 2r of every tile
  const long long am = m0 + a_row;
  int ab = 0, at = 0;
  const bool a_ok = am < M;
  if (a_ok) { ab = (int)(am / p.T); at = (int)(am - (long long)ab * p.T); }
  const float* xrow = p.x + (size_t)ab * p.x_bstride;
  const int half = p.ks / 2;
  const int b_n = tid / FBK, b_k = tid % FBK;                // weights: rows b_n + 16 r', column b_k of the tile
  float ra[8], rb[4], rg[EPI == 2 ? 4 : 1];
  auto fetch = [&](int k0) {
    const int j0 = k0 / p.Cin, j1 = min(k0 + FBK - 1, KD - 1) / p.Cin;      // block-uniform
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int kd = k0 + a_k + 2 * r;
      float v = 0.f;
      if (a_ok && kd < KD) {
        const int j = j0 == j1 ? j0 : kd / p.Cin;
        const int ci = kd - j * p.Cin;
        const int ti = at + p.dil * (j - half);
        v = (ti >= 0 && ti < p.T) ? __ldg(xrow + (size_t)ci * p.T + ti) : p.pad_value;
      }
      ra[r] = v;
    }
    const int kd = k0 + b_k;
    const int j = j0 == j1 ? j0 : kd / p.Cin;
    const int wcol = (kd - j * p.Cin) * p.ks + j;            // w[n][ci][j]
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int n = n0 + b_n + 16 * r;
      const bool ok = n < p.N && kd < KD;
      rb[r] = ok ? __ldg(p.w + (size_t)n * KD + wcol) : 0.f;
      if (EPI == 2) rg[r] = ok ? __ldg(p.w + (size_t)(n + p.N) * KD + wcol) : 0.f;
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int r = 0; r < 8; ++r) As[a_k + 2 * r][a_row] = ra[r];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      Bs[b_k][b_n + 16 * r] = rb[r];
      if (EPI == 2) Bg[b_k][b_n + 16 * r] = rg[r];
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < KD; k0 += FBK) {
    stage();
    __syncthreads();
    if (k0 + FBK < KD) fetch(k0 + FBK);
#pragma unroll
    for (int kk = 0; kk < FBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tx * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + tx * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (EPI == 2) {
        const float4 g4 = *reinterpret_cast<const float4*>(&Bg[kk][ty * 4]);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acg[i][j] = fmaf(a[i], g[j], acg[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
    if (m >= M) continue;
    const int b = (int)(m / p.T), t = (int)(m - (long long)b * p.T);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + ty * 4 + j;
      if (n >= p.N) continue;
      float* yo = p.y + (size_t)b * p.y_bstride + (size_t)n * p.T + t;
      float v = acc[i][j] + (p.bias ? __ldg(p.bias + n) : 0.f);
      if (EPI == 0) *yo = v;
      else if (EPI == 1) *yo += v;
      else {
        const float* ad = p.add + (size_t)b * p.add_bstride + t;
        const float pa = v + __ldg(ad + (size_t)n * p.T);
        const float pb = acg[i][j] + __ldg(p.bias + n + p.N) + __ldg(ad + (size_t)(n + p.N) * p.T);
        *yo = gated_unit(p.gate, pa, pb);                               // GTU: glow.py:33-40
      }
    }
  }
}

// z1 = (z1 - t) / exp(log_s), e = [log_s | t] (WN returns end(output).chunk(2, 1); modules.py:43-46)
__global__ void k_fd_coupling(float* __restrict__ z1, long long z_bstride, const float* __restrict__ e, int B, int n_half, int T,
                              int ignore_nan = 0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * n_half * T) return;
  const int t = (int)(i % T); const long long bc = i / T;
  const int c = (int)(bc % n_half), b = (int)(bc / n_half);
  const float log_s = e[((size_t)b * 2 * n_half + c) * T + t], tt = e[((size_t)b * 2 * n_half + n_half + c) * T + t];
  float* zp = z1 + (size_t)b * z_bstride + (size_t)c * T + t;
  float v = (*zp - tt) / expf(log_s);
  if (ignore_nan && v != v) v = 0.f;                                 // efficient_model_ax.py:331-332
  *zp = v;
}

// z.view(B, -1, G).transpose(1, 2) and back: 32 x 32 tiles through shared memory
__global__ void k_axg_transpose(const float* __restrict__ in, float* __restrict__ out, int T, int G, int to_cf) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, g0 = blockIdx.y * 32;
  const float* ib = in + (size_t)b * T * G; float* ob = out + (size_t)b * T * G;
  // channels-last element (t, g) = [t * G + g]; channels-first = [g * T + t]
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    if (to_cf) { const int t = t0 + r, g = g0 + threadIdx.x; if (t < T && g < G) tile[r][threadIdx.x] = ib[(size_t)t * G + g]; }
    else       { const int g = g0 + r, t = t0 + threadIdx.x; if (t < T && g < G) tile[r][threadIdx.x] = ib[(size_t)g * T + t]; }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    if (to_cf) { const int g = g0 + r, t = t0 + threadIdx.x; if (t < T && g < G) ob[(size_t)g * T + t] = tile[threadIdx.x][r]; }
    else       { const int t = t0 + r, g = g0 + threadIdx.x; if (t < T && g < G) ob[(size_t)t * G + g] = tile[threadIdx.x][r]; }
  }
}

__global__ void k_fd_add(float* __restrict__ y, const float* __restrict__ x, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}

template <int EPI>
int conv(const FdConvP& p, cudaStream_t s) {
  const long long M = (long long)p.B * p.T;
  dim3 grid((unsigned)((M + FBM - 1) / FBM), (unsigned)((p.N + FBN - 1) / FBN));
  k_fd_conv<EPI><<<grid, 256, 0, s>>>(p);
  CWG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int check(const cwg_fd_config* c, int batch, int T) {
  CWG_REQUIRE(c != nullptr, "cfg is NULL");
  CWG_REQUIRE(c->n_group >= 2 && c->n_group % 2 == 0 && c->n_flows >= 1 && c->n_early_every >= 1 && c->n_early_size % 2 == 0,
              "bad n_group / n_flows / early-output settings");
  CWG_REQUIRE(c->n_layers >= 1 && c->n_layers <= CWG_FD_MAX_LAYERS && c->n_channels >= 1 && c->kernel_size % 2 == 1 && c->cond_channels >= 1,
              "bad WN settings");
  CWG_REQUIRE(c->res_skip || c->merge_res_skip, "cannot remove res_skip without merge_res_skip (glow.py:53)");
  int n_rem = c->n_group;
  for (int k = 1; k < c->n_flows; ++k) if (k % c->n_early_every == 0) n_rem -= c->n_early_size;
  CWG_REQUIRE(n_rem >= 2, "too many early outputs for n_group");
  for (int i = 0; i < c->n_layers; ++i) CWG_REQUIRE(c->dilations[i] >= 1, "dilations must be >= 1");
  CWG_REQUIRE(batch >= 1 && T >= 1, "batch and T must be >= 1");
  return 0;
}

struct FdWs { float *z2, *h, *out, *acts, *c_all, *e; size_t bytes; };
void carve(const cwg_fd_config* c, int B, int T, void* base, FdWs* ws) {
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n * sizeof(float), 256); return (float*)((char*)base + o); };
  const size_t BT = (size_t)B * T;
  ws->z2 = take(BT * c->n_group); ws->h = take(BT * c->n_channels); ws->out = take(BT * c->n_channels);
  ws->acts = take(BT * c->n_channels); ws->c_all = take(BT * 2 * c->n_channels * c->n_layers); ws->e = take(BT * c->n_group);
  ws->bytes = off;
}

int axg_check(const cwg_axg_config* c, int batch, int T) {
  CWG_REQUIRE(c != nullptr, "cfg is NULL");
  CWG_REQUIRE(c->n_group >= 2 && c->n_rem >= 2 && c->n_rem % 2 == 0 && c->n_rem <= c->n_group, "bad n_group / n_rem");
  CWG_REQUIRE(c->n_layers >= 1 && c->n_layers <= CWG_FD_MAX_LAYERS && c->n_channels >= 1 && c->kernel_size % 2 == 1, "bad WN settings");
  CWG_REQUIRE(c->res_skip || c->merge_res_skip || c->n_layers == 1, "cannot remove res_skip without merge_res_skip (glow_ax.py:259)");
  CWG_REQUIRE(c->gate >= 0 && c->gate < CWG_GATE_COUNT, "unknown gated unit %d", c->gate);
  for (int i = 0; i < c->n_layers; ++i) CWG_REQUIRE(c->dilations[i] >= 1, "dilations must be >= 1");
  CWG_REQUIRE(batch >= 1 && T >= 1, "batch and T must be >= 1");
  return 0;
}

struct AxgWs { float *z2, *h, *out, *acts, *e; size_t bytes; };
void axg_carve(const cwg_axg_config* c, int B, int T, void* base, AxgWs* ws) {
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n * sizeof(float), 256); return (float*)((char*)base + o); };
  const size_t BT = (size_t)B * T;
  ws->z2 = take(BT * c->n_group); ws->h = take(BT * c->n_channels); ws->out = take(BT * c->n_channels);
  ws->acts = take(BT * c->n_channels); ws->e = take(BT * c->n_group);
  ws->bytes = off;
}

}  // namespace
}  // namespace cwg

using namespace cwg;

extern "C" {

size_t cwg_fd_workspace_bytes(const cwg_fd_config* cfg, int batch, int t_steps) {
  if (check(cfg, batch, t_steps)) return 0;
  FdWs ws;
  carve(cfg, batch, t_steps, nullptr, &ws);
  return ws.bytes;
}

int cwg_fd_launch_count(const cwg_fd_config* cfg) {
  if (check(cfg, 1, 1)) return -1;
  const int per_layer = 1 + (cfg->res_skip ? (cfg->merge_res_skip ? 1 : 2) : 1);
  return 1 + cfg->n_flows * (2 + cfg->n_layers * per_layer + 4);
}

int cwg_fd_inverse(const cwg_fd_config* cfg, const cwg_fd_weights* w, const float* cond, float* z,
                   void* workspace, size_t workspace_bytes, int batch, int t_steps, void* cuda_stream) {
  if (int r = check(cfg, batch, t_steps)) return r;
  CWG_REQUIRE(w && z && workspace && (cond || w->c_all), "NULL argument");
  CWG_REQUIRE(w->start_w && w->start_b && (w->c_all || (w->cond_w && w->cond_b)) && w->in_w && w->in_b && w->end_w && w->end_b && w->winv,
              "missing weight arrays");
  CWG_REQUIRE(!cfg->res_skip || (w->rs_w && w->rs_b), "res_skip weights missing");
  FdWs ws;
  carve(cfg, batch, t_steps, workspace, &ws);
  CWG_REQUIRE(ws.bytes <= workspace_bytes, "workspace too small: need %zu, got %zu", ws.bytes, workspace_bytes);
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const int B = batch, T = t_steps, G = cfg->n_group, C = cfg->n_channels, L = cfg->n_layers, ks = cfg->kernel_size, F = cfg->n_flows;
  const long long zb = (long long)G * T;
  // per-flow channel counts and offsets into the concatenated arrays
  int n_rem_k[256];
  size_t o_start[256], o_end_w[256], o_end_b[256], o_winv[256];
  CWG_REQUIRE(F <= 256, "n_flows > 256");
  {
    int n_rem = G; size_t a = 0, b = 0, c = 0, d = 0;
    for (int k = 0; k < F; ++k) {
      if (k % cfg->n_early_every == 0 && k > 0) n_rem -= cfg->n_early_size;
      n_rem_k[k] = n_rem;
      const int nh = n_rem / 2;
      o_start[k] = a; a += (size_t)C * nh;
      o_end_w[k] = b; b += (size_t)2 * nh * C;
      o_end_b[k] = c; c += (size_t)2 * nh;
      o_winv[k] = d; d += (size_t)n_rem * n_rem;
    }
  }
  // the early-output channels never change: both ping-pong buffers carry them from the start
  CWG_CHECK_CUDA(cudaMemcpyAsync(ws.z2, z, (size_t)B * G * T * sizeof(float), cudaMemcpyDeviceToDevice, s));
  float* cur = z; float* other = ws.z2;
  auto mix = [&](int k) -> int {                     // InvertibleConv1x1.inverse: conv1d with W^-1 over the active channels
    const int n_rem = n_rem_k[k], off = G - n_rem;
    FdConvP p{};
    p.B = B; p.Cin = n_rem; p.T = T; p.N = n_rem; p.ks = 1; p.dil = 1; p.pad_value = 0.f;
    p.x = cur + (size_t)off * T; p.x_bstride = zb; p.w = w->winv + o_winv[k]; p.bias = nullptr;
    p.y = other + (size_t)off * T; p.y_bstride = zb;
    if (int r = conv<0>(p, s)) return r;
    float* t = cur; cur = other; other = t;
    return 0;
  };
  for (int k = F - 1; k >= 0; --k) {
    const int n_rem = n_rem_k[k], nh = n_rem / 2, off = G - n_rem;
    if (!cfg->mix_first) { if (int r = mix(k)) return r; }
    // ---- WN(z_0, cond): start, cond layer, layers, end
    FdConvP p{};
    p.B = B; p.T = T; p.dil = 1; p.pad_value = 0.f;
    p.Cin = nh; p.N = C; p.ks = 1; p.x = cur + (size_t)off * T; p.x_bstride = zb;
    p.w = w->start_w + o_start[k]; p.bias = w->start_b + (size_t)k * C; p.y = ws.h; p.y_bstride = (long long)C * T;
    if (int r = conv<0>(p, s)) return r;
    const float* c_all = ws.c_all;
    if (w->c_all) {                                                                 // cond stack evaluated by the caller
      c_all = w->c_all + (size_t)k * B * 2 * C * L * T;
    } else {
      p.Cin = cfg->cond_channels; p.N = 2 * C * L; p.x = cond; p.x_bstride = (long long)cfg->cond_channels * T;
      p.w = w->cond_w + (size_t)k * 2 * C * L * cfg->cond_channels; p.bias = w->cond_b + (size_t)k * 2 * C * L;
      p.y = ws.c_all; p.y_bstride = (long long)2 * C * L * T;
      if (int r = conv<0>(p, s)) return r;
    }
    if (!cfg->merge_res_skip) CWG_CHECK_CUDA(cudaMemsetAsync(ws.out, 0, (size_t)B * C * T * sizeof(float), s));
    for (int i = 0; i < L; ++i) {
      const size_t li = (size_t)k * L + i;
      FdConvP g{};
      g.B = B; g.T = T; g.Cin = C; g.N = C; g.ks = ks; g.dil = cfg->dilations[i];
      g.pad_value = k == 0 ? cfg->first_pad_value : 0.f;                        // untts glow.py:80,126
      g.x = ws.h; g.x_bstride = (long long)C * T; g.w = w->in_w + li * 2 * C * C * ks; g.bias = w->in_b + li * 2 * C;
      g.add = c_all + (size_t)2 * C * i * T; g.add_bstride = (long long)2 * C * L * T;
      g.y = ws.acts; g.y_bstride = (long long)C * T;
      if (int r = conv<2>(g, s)) return r;
      const bool last = i == L - 1;
      if (cfg->res_skip) {
        FdConvP q{};
        q.B = B; q.T = T; q.Cin = C; q.N = C; q.ks = 1; q.dil = 1; q.x = ws.acts; q.x_bstride = (long long)C * T;
        q.w = w->rs_w + li * 2 * C * C; q.bias = w->rs_b + li * 2 * C; q.y_bstride = (long long)C * T;
        if (cfg->merge_res_skip) { q.y = ws.h; if (int r = conv<1>(q, s)) return r; }
        else if (!last) {
          q.y = ws.h; if (int r = conv<1>(q, s)) return r;                        // res rows [0, C)
          q.w += (size_t)C * C; q.bias += C; q.y = ws.out; if (int r = conv<1>(q, s)) return r;   // skip rows [C, 2C)
        } else { q.y = ws.out; if (int r = conv<1>(q, s)) return r; }
      } else {
        // no res_skip layer: res_skip_acts = acts and (merged) h += acts (glow.py:149-155)
        const long long n = (long long)B * C * T;
        k_fd_add<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ws.h, ws.acts, n);
        CWG_CHECK_CUDA(cudaGetLastError());
      }
    }
    FdConvP e{};
    e.B = B; e.T = T; e.Cin = C; e.N = 2 * nh; e.ks = 1; e.dil = 1;
    e.x = cfg->merge_res_skip ? ws.h : ws.out; e.x_bstride = (long long)C * T;
    e.w = w->end_w + o_end_w[k]; e.bias = w->end_b + o_end_b[k]; e.y = ws.e; e.y_bstride = (long long)2 * nh * T;
    if (int r = conv<0>(e, s)) return r;
    {
      const long long n = (long long)B * nh * T;
      k_fd_coupling<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cur + (size_t)(off + nh) * T, zb, ws.e, B, nh, T);
      CWG_CHECK_CUDA(cudaGetLastError());
    }
    if (cfg->mix_first) { if (int r = mix(k)) return r; }
  }
  if (cur != z) CWG_CHECK_CUDA(cudaMemcpyAsync(z, cur, (size_t)B * G * T * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}

size_t cwg_axg_workspace_bytes(const cwg_axg_config* cfg, int batch, int t_steps) {
  if (axg_check(cfg, batch, t_steps)) return 0;
  AxgWs ws;
  axg_carve(cfg, batch, t_steps, nullptr, &ws);
  return ws.bytes;
}

int cwg_axg_launch_count(const cwg_axg_config* cfg) {
  if (axg_check(cfg, 1, 1)) return -1;
  const bool split = cfg->res_skip && !cfg->merge_res_skip;
  return 2 /* mix + copy-back */ + 2 /* start, memset */ + cfg->n_layers * 2 + (split ? cfg->n_layers - 1 : 0) + 2 /* end, coupling */;
}

// One flow of efficient_model_ax.py:325-340 with glow_ax.WN.forward :375-418 in the general form.
int cwg_axg_flow(const cwg_axg_config* cfg, const cwg_axg_weights* w, const float* c_all, float* z,
                 void* workspace, size_t workspace_bytes, int batch, int t_steps, void* cuda_stream) {
  if (int r = axg_check(cfg, batch, t_steps)) return r;
  CWG_REQUIRE(w && c_all && z && workspace, "NULL argument");
  CWG_REQUIRE(w->start_w && w->start_b && w->in_w && w->in_b && w->end_w && w->end_b && w->winv, "missing weight arrays");
  CWG_REQUIRE(!cfg->res_skip || (w->rs_w && w->rs_b), "res_skip weights missing");
  AxgWs ws;
  axg_carve(cfg, batch, t_steps, workspace, &ws);
  CWG_REQUIRE(ws.bytes <= workspace_bytes, "workspace too small: need %zu, got %zu", ws.bytes, workspace_bytes);
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const int B = batch, T = t_steps, G = cfg->n_group, C = cfg->n_channels, L = cfg->n_layers, ks = cfg->kernel_size;
  const int n_rem = cfg->n_rem, nh = n_rem / 2, off = G - n_rem;
  const long long zb = (long long)G * T, cb = (long long)C * T;
  auto mix = [&]() -> int {                          // InvertibleConv1x1.inverse / PermuteHeight.inverse over the active rows
    FdConvP p{};
    p.B = B; p.Cin = n_rem; p.T = T; p.N = n_rem; p.ks = 1; p.dil = 1;
    p.x = z + (size_t)off * T; p.x_bstride = zb; p.w = w->winv; p.bias = nullptr;
    p.y = ws.z2 + (size_t)off * T; p.y_bstride = zb;
    if (int r = conv<0>(p, s)) return r;
    CWG_CHECK_CUDA(cudaMemcpy2DAsync(z + (size_t)off * T, (size_t)zb * sizeof(float), ws.z2 + (size_t)off * T, (size_t)zb * sizeof(float),
                                     (size_t)n_rem * T * sizeof(float), B, cudaMemcpyDeviceToDevice, s));
    return 0;
  };
  if (!cfg->mix_first) { if (int r = mix()) return r; }
  FdConvP p{};
  p.B = B; p.T = T; p.dil = 1; p.Cin = nh; p.N = C; p.ks = 1; p.x = z + (size_t)off * T; p.x_bstride = zb;
  p.w = w->start_w; p.bias = w->start_b; p.y = ws.h; p.y_bstride = cb;
  if (int r = conv<0>(p, s)) return r;                                              // glow_ax.py:376
  CWG_CHECK_CUDA(cudaMemsetAsync(ws.out, 0, (size_t)B * C * T * sizeof(float), s));
  const bool split = cfg->res_skip && !cfg->merge_res_skip;
  for (int i = 0; i < L; ++i) {
    FdConvP g{};
    g.B = B; g.T = T; g.Cin = C; g.N = C; g.ks = ks; g.dil = cfg->dilations[i]; g.gate = cfg->gate;
    g.x = ws.h; g.x_bstride = cb; g.w = w->in_w + (size_t)i * 2 * C * C * ks; g.bias = w->in_b + (size_t)i * 2 * C;
    g.add = c_all + (size_t)2 * C * i * T; g.add_bstride = (long long)2 * C * L * T;
    g.y = ws.acts; g.y_bstride = cb;
    if (int r = conv<2>(g, s)) return r;                                            // :394-397
    if (cfg->res_skip) {
      FdConvP q{};
      q.B = B; q.T = T; q.Cin = C; q.N = C; q.ks = 1; q.dil = 1; q.x = ws.acts; q.x_bstride = cb;
      q.w = w->rs_w + (size_t)i * 2 * C * C; q.bias = w->rs_b + (size_t)i * 2 * C; q.y_bstride = cb;
      if (split && i < L - 1) {                                                     // :402-404, :409-411
        q.y = ws.h; if (int r = conv<1>(q, s)) return r;
        q.w += (size_t)C * C; q.bias += C;
      }
      q.y = ws.out; if (int r = conv<1>(q, s)) return r;                            // :406, :413 (merged: the hidden tensor stays)
    } else {
      const long long n = (long long)B * C * T;                                     // :399: res_skip_acts = acts
      k_fd_add<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ws.out, ws.acts, n);
      CWG_CHECK_CUDA(cudaGetLastError());
    }
  }
  FdConvP e{};
  e.B = B; e.T = T; e.Cin = C; e.N = 2 * nh; e.ks = 1; e.dil = 1; e.x = ws.out; e.x_bstride = cb;
  e.w = w->end_w; e.bias = w->end_b; e.y = ws.e; e.y_bstride = (long long)2 * nh * T;
  if (int r = conv<0>(e, s)) return r;                                              // :418
  {
    const long long n = (long long)B * nh * T;
    k_fd_coupling<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(z + (size_t)(off + nh) * T, zb, ws.e, B, nh, T, cfg->ignore_nan);
    CWG_CHECK_CUDA(cudaGetLastError());
  }
  if (cfg->mix_first) { if (int r = mix()) return r; }
  return 0;
}

int cwg_group_transpose(const float* in, float* out, int batch, int t_steps, int n_group, int to_channels_first, void* cuda_stream) {
  CWG_REQUIRE(in && out && in != out, "in / out must be distinct device buffers");
  CWG_REQUIRE(batch >= 1 && batch <= 65535 && t_steps >= 1 && n_group >= 1, "bad shape");
  dim3 grid((unsigned)((t_steps + 31) / 32), (unsigned)((n_group + 31) / 32), (unsigned)batch);
  k_axg_transpose<<<grid, dim3(32, 8), 0, (cudaStream_t)cuda_stream>>>(in, out, t_steps, n_group, to_channels_first);
  CWG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"

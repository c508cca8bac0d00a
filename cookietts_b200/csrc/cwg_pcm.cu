// 16-bit PCM conversion of the vocoder output: what the reference's T2S server does on the host
// per utterance after every vocoder call (CookieTTS/_5_infer/t2s_server/text2speech.py:672-694): trim to
// output_length*hop samples, append `cat_silence_s` of zeros, `(audio * 2**15).astype('int16')`.
// HBM-bound: reads 4 B, writes 2 B per sample; 128-bit accesses (8 samples per thread).
#include "cwg_common.cuh"

namespace cwg {
namespace {

// numpy's float32 -> int16 cast: truncate toward zero.  Out-of-range values are undefined behaviour in C; the
// scalar x86 path numpy compiles to (cvttss2si, keep the low 16 bits) is reproduced for `saturate == 0`.
__device__ __forceinline__ short to_pcm16(float v, int saturate) {
  const float s = v * 32768.0f;
  if (saturate) return (short)max(-32768, min(32767, __float2int_rz(s)));   // NaN -> 0
  if (!(fabsf(s) < 2147483648.0f)) return 0;                                 // "integer indefinite" 0x80000000
  return (short)(unsigned short)(__float2int_rz(s) & 0xffff);
}

__global__ void __launch_bounds__(256) k_pcm16(const float* __restrict__ audio, int t_stride,
                                               const int* __restrict__ n_valid, short* __restrict__ out,
                                               int out_stride, int saturate, int vec_ok) {
  const int b = blockIdx.y;
  const int valid = n_valid ? min(max(n_valid[b], 0), t_stride) : t_stride;
  const float* src = audio + (size_t)b * t_stride;
  short* dst = out + (size_t)b * out_stride;
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i0 >= out_stride) return;
  if (vec_ok && i0 + 8 <= out_stride) {
    alignas(16) short r[8];
    if (i0 + 8 <= valid) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(src + i0));
      const float4 c = __ldcs(reinterpret_cast<const float4*>(src + i0 + 4));
      r[0] = to_pcm16(a.x, saturate); r[1] = to_pcm16(a.y, saturate); r[2] = to_pcm16(a.z, saturate); r[3] = to_pcm16(a.w, saturate);
      r[4] = to_pcm16(c.x, saturate); r[5] = to_pcm16(c.y, saturate); r[6] = to_pcm16(c.z, saturate); r[7] = to_pcm16(c.w, saturate);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = (i0 + j < valid) ? to_pcm16(src[i0 + j], saturate) : (short)0;
    }
    __stcs(reinterpret_cast<int4*>(dst + i0), *reinterpret_cast<const int4*>(r));
  } else {
    for (int j = 0; j < 8 && i0 + j < out_stride; ++j)
      dst[i0 + j] = (i0 + j < valid) ? to_pcm16(src[i0 + j], saturate) : (short)0;
  }
}

}  // namespace
}  // namespace cwg

using namespace cwg;

extern "C" int cwg_pcm16(const float* audio, int batch, int t_stride, const int32_t* n_valid, int16_t* out,
                         int out_stride, int saturate, void* cuda_stream) {
  CWG_REQUIRE(batch >= 0 && t_stride >= 0 && out_stride >= 0, "negative size");
  if (batch == 0 || out_stride == 0) return 0;
  CWG_REQUIRE(audio && out, "NULL argument");
  const int vec_ok = (t_stride % 4 == 0) && (out_stride % 8 == 0) &&
                     ((uintptr_t)audio % 16 == 0) && ((uintptr_t)out % 16 == 0);
  dim3 grid((out_stride + 2047) / 2048, batch);
  k_pcm16<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(audio, t_stride, (const int*)n_valid, (short*)out, out_stride,
                                                       saturate, vec_ok);
  CWG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

"""Stage timings of the "next" rows at BASELINE config-2 size (16 utterances x 10 s @ 22.05 kHz) on one GPU:
Denoiser (STFT -> bias subtraction -> iSTFT), int16 PCM conversion, de-emphasis, and the ax conditioning front-end
(speaker embedding + 2 cond layers + 3 transposed convs, the `axfe_256` shape).  CUDA events, 3 warm-ups, median of 10;
prints one JSON line per stage with the algorithmic work and the achieved rate."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cookietts_b200 import Denoiser, serving, _cabi  # noqa: E402
from cookietts_b200.ax_frontend import repack_conv_transpose  # noqa: E402

B, T, SR = 16, 861 * 256, 22050
dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6538.9}


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


class _Voc(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1, device=dev))

    def infer(self, mel, speaker_ids=None, sigma=1.0):
        return torch.randn(1, mel.shape[2] * 256, device=dev) * 0.01


audio = torch.randn(B, T, device=dev) * 0.3
den = Denoiser(_Voc(), sampling_rate=SR, n_mel_channels=80)
fl, hop = den.filter_length, den.hop_length
nf = (T + 2 * (fl // 2) - fl) // hop + 1
ms = timed(lambda: den(audio, strength=0.1))
flops = 2.0 * 2 * B * nf * fl * 2 * den.cutoff                        # analysis + synthesis GEMMs
print(json.dumps({"stage": "denoiser", "ms": ms, "samples_per_s": B * T / ms * 1e3, "gemm_tflops": flops / ms / 1e9,
                  "filter_length": fl, "hop": hop, "frames": nf, "note": "fp32 CUDA-core GEMMs against the windowed Fourier bases"}))

for bb in (B, 256):                                                    # config 2 and config 3 batch sizes
    big = torch.randn(bb, T, device=dev) * 0.3
    nv = torch.full((bb,), T, dtype=torch.int32, device=dev)
    ms = timed(lambda: serving.pcm16(big, nv, 0))
    gb = bb * T * 6 / 1e9
    print(json.dumps({"stage": f"pcm16 (B={bb})", "ms": ms, "GBps": gb / ms * 1e3, "frac_of_hbm_peak": gb / ms * 1e3 / peaks["hbm_gbs"],
                      "bytes_per_sample": 6}))
    del big

lib = _cabi.load()
out = torch.empty_like(audio)
st = torch.cuda.current_stream().cuda_stream
ms = timed(lambda: _cabi.check(lib.cwg_deemphasis(audio.data_ptr(), B, T, 0.97, 0, out.data_ptr(), st)))
gb = B * T * 8 / 1e9
print(json.dumps({"stage": "deemphasis", "ms": ms, "GBps": gb / ms * 1e3, "note": "one CTA per utterance, fp64 chunked scan"}))

# ax front-end of the axfe_256 golden shape at 861 frames: 96 -> 64 -> 48 cond convs, then 48 -> 64 -> 64 -> 96 transposed convs x2 x4 x4
frames = 862
x = torch.randn(B, 96, frames, device=dev)
w1 = torch.randn(64, 96, 1, device=dev) * 0.1; b1 = torch.zeros(64, device=dev)
w2 = torch.randn(48, 64, 1, device=dev) * 0.1; b2 = torch.zeros(48, device=dev)
tws = []
cin = 48
for (cout, k, s) in ((64, 4, 2), (64, 8, 4), (96, 4, 4)):
    w = (torch.randn(cin, cout, k) * 0.1).numpy()
    tws.append((torch.from_numpy(repack_conv_transpose(w, s)).to(dev), torch.zeros(cout, device=dev), cin, cout, k, s, (k - s) // 2))
    cin = cout


def frontend():
    h1 = torch.empty(B, 64, frames, device=dev); h2 = torch.empty(B, 48, frames, device=dev)
    _cabi.check(lib.cwg_conv1d(x.data_ptr(), B, 96, frames, w1.data_ptr(), b1.data_ptr(), 64, 1, 0, 0, 2, 0.1, 1.0, None, h1.data_ptr(), st))
    _cabi.check(lib.cwg_conv1d(h1.data_ptr(), B, 64, frames, w2.data_ptr(), b2.data_ptr(), 48, 1, 0, 0, 2, 0.1, 1.0, None, h2.data_ptr(), st))
    h, t = h2, frames
    for (w, b, ci, co, k, s, p) in tws:
        to = (t - 1) * s - 2 * p + k
        y = torch.empty(B, co, to, device=dev)
        _cabi.check(lib.cwg_conv_transpose1d(h.data_ptr(), B, ci, t, w.data_ptr(), b.data_ptr(), co, k, s, p, 2, 0.4, 1.0, y.data_ptr(), st))
        h, t = y, to
    return h


ms = timed(frontend)
print(json.dumps({"stage": "ax_frontend", "ms": ms, "out_shape": list(frontend().shape)}))

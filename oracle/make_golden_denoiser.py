"""Golden vectors for the Denoiser post-filter from the UNMODIFIED reference
(`CookieTTS/utils/audio/stft.py::STFT` and `CookieTTS/_4_mtw/waveglow/denoiser.py::Denoiser.forward`).

    python oracle/make_golden_denoiser.py            (build container only)

librosa is not installed here; stft.py imports three names from it.  `pad_center` (zero-pad to a
centred length) and `tiny` (smallest normal float) are stubbed with their one-line definitions,
`librosa.filters.mel` is unused on this path.  Denoiser.__init__ needs a vocoder whose `infer` takes
`speaker_ids=` (denoiser.py:39-50), so the object is built with a tiny stand-in vocoder returning a
fixed "bias audio"; `forward` is then the reference's own code.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def install_stubs():
    lib = types.ModuleType("librosa"); util = types.ModuleType("librosa.util"); filt = types.ModuleType("librosa.filters")

    def pad_center(data, size, axis=-1, **kw):
        n = data.shape[axis]; lpad = int((size - n) // 2)
        lengths = [(0, 0)] * data.ndim; lengths[axis] = (lpad, int(size - n - lpad))
        return np.pad(data, lengths)

    def tiny(x):
        x = np.asarray(x)
        return np.finfo(x.dtype if np.issubdtype(x.dtype, np.floating) else np.float32).tiny

    util.pad_center, util.tiny, util.normalize = pad_center, tiny, (lambda x, norm=None, **k: x)
    filt.mel = lambda *a, **k: None
    lib.util, lib.filters = util, filt
    sys.modules.update({"librosa": lib, "librosa.util": util, "librosa.filters": filt})
    iso = types.ModuleType("CookieTTS.utils.audio.iso226"); iso.ISO_226 = object
    sys.modules["CookieTTS.utils.audio.iso226"] = iso
    sys.path.insert(0, "/root/reference")


class FakeVocoder(torch.nn.Module):
    def __init__(self, bias_audio):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.bias_audio = bias_audio

    def infer(self, mel, speaker_ids=None, sigma=1.0):
        return self.bias_audio.clone()


CASES = {  # name: (sampling_rate, batch, n_samples, strength, seed)
    "denoiser_22k": (22050, 2, 6000, 0.1, 1),
    "denoiser_48k": (48000, 1, 9000, 0.35, 2),
    "denoiser_small": (4000, 3, 1234, 1.5, 3),
}


def main():
    install_stubs()
    from CookieTTS._4_mtw.waveglow.denoiser import Denoiser
    outdir = os.path.join(ROOT, "tests", "golden")
    for name, (sr, B, T, strength, seed) in CASES.items():
        rs = np.random.RandomState(seed)
        audio = (rs.standard_normal((B, T)) * 0.3).astype(np.float32)
        bias_audio = (rs.standard_normal((1, 20 * (sr // 400) * 4)) * 0.01).astype(np.float32)
        den = Denoiser(FakeVocoder(torch.from_numpy(bias_audio)), sampling_rate=sr, n_mel_channels=80)
        with torch.no_grad():
            out = den(torch.from_numpy(audio), strength=strength)
            mag, _ = den.stft.transform(torch.from_numpy(audio))
        print(name, "fl/hop", den.stft.filter_length, den.stft.hop_length, "out", tuple(out.shape),
              "bias", tuple(den.bias_spec.shape), "delta", float((out[:, 0, :] - torch.from_numpy(audio)[:, :out.shape[2]]).abs().max()))
        np.savez_compressed(os.path.join(outdir, name + ".npz"), sampling_rate=sr, strength=strength, audio=audio,
                            bias_audio=bias_audio, bias_spec=den.bias_spec.numpy(), magnitude=mag.numpy(),
                            denoised=out.numpy())


if __name__ == "__main__":
    main()

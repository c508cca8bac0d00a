"""GPU parity of the ax models with the conditioning front-end / output filters switched on against
golden vectors of the unmodified reference (tests/golden/axfe_*.npz), and of the fp32 building-block kernels
(csrc/cwg_condnet.cu) against plain torch fp32 references of the same ops."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cookietts_b200 import WaveGlowAx, WaveFlow, _cabi
from cookietts_b200.ax_frontend import repack_conv_transpose
from oracle.waveglow_oracle import snr_db
from tests.ax_frontend_helpers import CASES, load_case, module_kwargs
from tests.helpers import max_abs

pytestmark = pytest.mark.gpu

TOL = {"ffma": dict(max_abs=1e-4, snr=100.0), "bf16x3": dict(max_abs=1e-3, snr=60.0), "bf16": dict(max_abs=5e-2, snr=40.0),
       "f16f8": dict(max_abs=1e-3, snr=60.0)}


def run(name, precision):
    kind, cfg, fe, sd, g = load_case(name)
    m = (WaveGlowAx if kind == "ax" else WaveFlow)(precision=precision, **module_kwargs(kind, cfg, fe))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    m = m.cuda().eval()
    aud = m.infer(torch.from_numpy(g["mel"]).cuda(), speaker_ids=torch.from_numpy(g["speaker_ids"]).cuda(),
                  sigma=float(g["sigma"]), z=torch.from_numpy(g["z"]).cuda())
    return aud.numpy(), g["infer_ref_fp64"]


def check(out, ref, precision):
    assert out.shape == ref.shape and np.isfinite(out).all()
    scale = max(1.0, float(np.abs(ref).max()) / 4.0)                 # cases with the volume map reach |x| ~ 13
    assert max_abs(out, ref) <= TOL[precision]["max_abs"] * scale
    assert snr_db(ref, out) >= TOL[precision]["snr"]


@pytest.mark.parametrize("name", [c for c in CASES if "waveflow" not in c])
def test_frontend_fp32_cuda_cores(name):
    out, ref = run(name, "ffma")
    check(out, ref, "ffma")


@pytest.mark.parametrize("name,precision", [("axfe_256", "bf16x3"), ("axfe_256", "bf16"),
                                            ("axfe_waveflow", "bf16x3"), ("axfe_waveflow", "bf16"),
                                            ("axfe_separable_256", "bf16x3"), ("axfe_waveflow_separable", "bf16x3"),
                                            ("axfe_256", "f16f8"), ("axfe_separable_256", "f16f8"),
                                            # the notebook layout: n_group 24 (wide group padding), WN-level speaker
                                            # embeddings (two utterances, two speakers), upsample_first=False
                                            ("axfe_nb_256", "bf16x3"), ("axfe_nb_256", "bf16"), ("axfe_nb_256", "f16f8")])
def test_frontend_tensor_cores(name, precision):
    out, ref = run(name, precision)
    check(out, ref, precision)


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3"])
def test_notebook_model(precision):
    """The one configuration the reference records a speed for (scripts/WaveGlowFlow Inference Speed Testing.ipynb
    cell 2: 48 flows, n_group 24, 8 x 256 WN with 96-dim speaker embeddings, upsample_first=False) against the reference's
    own output for a seeded synthetic checkpoint (oracle/make_golden_ax_frontend.py, case axfe_notebook)."""
    out, ref = run("axfe_notebook", precision)
    check(out, ref, precision)


def test_wn_speaker_ids_are_per_utterance():
    """Each utterance gets its own speaker's gate bias: a batch equals its utterances run one at a time."""
    kind, cfg, fe, sd, g = load_case("axfe_nb_256")
    m = WaveGlowAx(precision="bf16x3", **module_kwargs(kind, cfg, fe))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    m = m.cuda().eval()
    mel, z, spk = torch.from_numpy(g["mel"]).cuda(), torch.from_numpy(g["z"]).cuda(), torch.from_numpy(g["speaker_ids"]).cuda()
    both = m.infer(mel, speaker_ids=spk, sigma=0.9, z=z)
    for b in range(mel.shape[0]):
        one = m.infer(mel[b:b + 1], speaker_ids=spk[b:b + 1], sigma=0.9, z=z[b:b + 1])
        assert torch.equal(one[0], both[b])
    swapped = m.infer(mel, speaker_ids=spk.flip(0), sigma=0.9, z=z)
    assert not torch.allclose(swapped, both)


def test_missing_speaker_ids_raise():
    kind, cfg, fe, sd, g = load_case("axfe_speaker_cond")
    m = WaveGlowAx(precision="ffma", **module_kwargs(kind, cfg, fe)).cuda().eval()
    with pytest.raises(Exception, match="requires speaker ids"):     # efficient_model_ax.py:288-289
        m.infer(torch.from_numpy(g["mel"]).cuda(), sigma=1.0)


# ------------------------------------------------------------------ building blocks vs torch fp32
def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("cin,cout,k,pad,mode,act", [(7, 70, 3, 1, "zeros", 0), (33, 5, 5, 2, "replicate", 2), (16, 16, 1, 0, "zeros", 3),
                                                      (9, 65, 3, 1, "reflect", 1), (4, 4, 3, 1, "circular", 4)])
def test_conv1d_kernel(cin, cout, k, pad, mode, act):
    lib = _cabi.load()
    g = torch.Generator().manual_seed(cin * 100 + k)
    x = torch.randn(3, cin, 131, generator=g); w = torch.randn(cout, cin, k, generator=g) * 0.2; b = torch.randn(cout, generator=g)
    res = torch.randn(3, cout, 131 + 2 * pad - (k - 1), generator=g)
    xp = F.pad(x, (pad, pad), mode={"zeros": "constant", "replicate": "replicate", "reflect": "reflect", "circular": "circular"}[mode]) if pad else x
    ref = F.conv1d(xp.double(), w.double(), b.double())
    ref = [lambda v: v, torch.relu, lambda v: F.leaky_relu(v, 0.3), torch.tanh, torch.sigmoid][act](ref)
    ref = res.double() + 0.7 * ref
    xd, wd, bd, rd = x.cuda(), w.cuda(), b.cuda(), res.cuda()
    y = torch.empty_like(rd)
    _cabi.check(lib.cwg_conv1d(xd.data_ptr(), 3, cin, 131, wd.data_ptr(), bd.data_ptr(), cout, k, pad,
                               {"zeros": 0, "replicate": 1, "reflect": 2, "circular": 3}[mode], act, 0.3, 0.7, rd.data_ptr(),
                               y.data_ptr(), _stream()))
    assert max_abs(y.cpu().numpy(), ref.numpy()) < 2e-5


@pytest.mark.parametrize("cin,cout,k,s", [(6, 70, 4, 2), (5, 9, 9, 3), (8, 8, 5, 5), (3, 4, 3, 2), (4, 3, 8, 4), (2, 2, 2, 3)])
def test_conv_transpose1d_kernel(cin, cout, k, s):
    lib = _cabi.load()
    g = torch.Generator().manual_seed(cin * 10 + k)
    p = max((k - s) // 2, 0)
    x = torch.randn(2, cin, 37, generator=g); w = torch.randn(cin, cout, k, generator=g) * 0.3; b = torch.randn(cout, generator=g)
    ref = F.leaky_relu(F.conv_transpose1d(x.double(), w.double(), b.double(), stride=s, padding=p), 0.4) * 1.5
    wp = torch.from_numpy(repack_conv_transpose(w.numpy(), s)).cuda()
    xd, bd = x.cuda(), b.cuda()
    y = torch.full(tuple(ref.shape), float("nan"), device="cuda")
    _cabi.check(lib.cwg_conv_transpose1d(xd.data_ptr(), 2, cin, 37, wp.data_ptr(), bd.data_ptr(), cout, k, s, p, 2, 0.4, 1.5,
                                         y.data_ptr(), _stream()))
    assert max_abs(y.cpu().numpy(), ref.numpy()) < 2e-5


@pytest.mark.parametrize("mode,tin,tout", [(0, 10, 37), (1, 10, 37), (2, 10, 40), (1, 9, 9), (0, 12, 5)])
def test_resample1d_kernel(mode, tin, tout):
    lib = _cabi.load()
    x = torch.randn(2, 5, tin, generator=torch.Generator().manual_seed(tin + tout))
    if mode == 0:
        ref, sf = F.interpolate(x, size=tout, mode="nearest"), 0.0
    elif mode == 1:
        ref, sf = F.interpolate(x, size=tout, mode="linear", align_corners=True), 0.0
    else:
        ref, sf = F.interpolate(x, scale_factor=tout // tin, mode="linear", align_corners=False), float(tout // tin)
    crop = 1 if tout > 6 else 0
    base = torch.randn(2, 7, tout - 2 * crop, generator=torch.Generator().manual_seed(1))
    y = base.clone().cuda()
    xd = x.cuda()
    _cabi.check(lib.cwg_resample1d(xd.data_ptr(), 2, 5, tin, 5 * tin, y.data_ptr(), tout - 2 * crop, 7 * (tout - 2 * crop), mode,
                                   tout, crop, sf, 1, _stream()))
    expect = base.clone()
    expect[:, :5] += ref[:, :, crop:tout - crop]
    assert max_abs(y.cpu().numpy(), expect.numpy()) < 1e-5


@pytest.mark.parametrize("T,coef,vol", [(5000, 0.97, 0), (1023, 0.9, 1), (220160, 0.97, 0), (7, 0.5, 1), (3000, 0.0, 1)])
def test_deemphasis_kernel(T, coef, vol):
    from scipy import signal
    lib = _cabi.load()
    x = (np.random.RandomState(T).standard_normal((3, T)) * 0.3).astype(np.float32)
    ref = x.astype(np.float64)
    if vol:
        t = torch.from_numpy(x.copy())
        t[t > 0] = 10 ** (t[t > 0].log2()); t[t < 0] = -(10 ** ((-t[t < 0]).log2()))
        ref = t.numpy().astype(np.float64)
    if coef:
        ref = np.stack([signal.lfilter([1], [1, -coef], r) for r in ref])
    xd = torch.from_numpy(x).cuda(); y = torch.empty_like(xd)
    _cabi.check(lib.cwg_deemphasis(xd.data_ptr(), 3, T, coef, vol, y.data_ptr(), _stream()))
    assert max_abs(y.cpu().numpy(), ref) < 2e-6 * max(1.0, np.abs(ref).max())


def test_ax_graph_replay_matches_eager():
    """graphs="auto": the second call of a shape is captured, later calls replay it on static buffers - same numbers as
    the eager launches, new inputs (mel, z, speaker ids) are honoured, and a weight edit drops the captured graphs."""
    kind, cfg, fe, sd, g = load_case("axfe_nb_256")
    kw = module_kwargs(kind, cfg, fe)
    eager = WaveGlowAx(precision="f16f8", graphs=False, **kw)
    auto = WaveGlowAx(precision="f16f8", **kw)
    for m in (eager, auto):
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        m.cuda().eval()
    mel, z, spk = torch.from_numpy(g["mel"]).cuda(), torch.from_numpy(g["z"]).cuda(), torch.from_numpy(g["speaker_ids"]).cuda()
    ref = eager.infer(mel, speaker_ids=spk, sigma=0.9, z=z)
    outs = [auto.infer(mel, speaker_ids=spk, sigma=0.9, z=z) for _ in range(4)]
    assert len(auto._graphs) == 1
    for o in outs:
        assert torch.equal(o, ref)
    mel2, z2, spk2 = mel * 0.5 - 1.0, z.flip(1).contiguous(), spk.flip(0).contiguous()
    assert torch.equal(auto.infer(mel2, speaker_ids=spk2, sigma=0.9, z=z2), eager.infer(mel2, speaker_ids=spk2, sigma=0.9, z=z2))
    with torch.no_grad():
        for m in (eager, auto):
            m.WN[0].WN.end.bias.add_(0.05)
    out3 = auto.infer(mel, speaker_ids=spk, sigma=0.9, z=z)
    assert torch.equal(out3, eager.infer(mel, speaker_ids=spk, sigma=0.9, z=z)) and not torch.equal(out3, ref)

"""Shared helpers for the tests: golden loading and error metrics."""
import json
import os

import numpy as np

from oracle.waveglow_oracle import OracleConfig, synthetic_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = OracleConfig(**json.loads(str(g["config"])))
    sd = synthetic_state_dict(cfg, int(g["weight_seed"]))
    return cfg, sd, g


def load_golden_regen(name):
    """Full-length goldens (oracle/make_golden.py big) hold only the reference's outputs: mel and z are regenerated from
    the stored seeds and checked against the stored CRC32s."""
    import zlib
    from oracle.waveglow_oracle import synthetic_inputs
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = OracleConfig(**json.loads(str(g["config"])))
    sd = synthetic_state_dict(cfg, int(g["weight_seed"]))
    mel, z = synthetic_inputs(cfg, int(g["batch"]), int(g["t_mel"]), int(g["input_seed"]))
    assert zlib.crc32(np.ascontiguousarray(mel).tobytes()) == int(g["mel_crc32"]), "regenerated mel differs from the golden's"
    assert zlib.crc32(np.ascontiguousarray(z).tobytes()) == int(g["z_crc32"]), "regenerated z differs from the golden's"
    return cfg, sd, g, mel, z


def max_abs(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())

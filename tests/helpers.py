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


def max_abs(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())

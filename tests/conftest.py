import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "token-pruning-audio-transformer_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with `-m gpu` on the GPU box")


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def make_case(meta):
    """Regenerate (state-dict, input) of a golden config and check them against the stored digests."""
    import hashlib
    from oracle import weights
    if meta["variant"] == "audiomae":
        sd = weights.make_audiomae_state_dict(meta["num_classes"], meta["T"], meta["wseed"], meta["flavour"])
    else:
        sd = weights.make_ast_state_dict(meta["num_classes"], meta["T"], meta["wseed"], meta["flavour"])
    x = weights.make_spectrogram(meta["variant"], meta["B"], meta["T"], meta["xseed"])
    assert weights.state_dict_digest(sd) == meta["sd_digest"], "weight RNG stream drifted from the golden generator"
    assert hashlib.sha256(x.numpy().tobytes()).hexdigest() == meta["x_digest"], "input RNG stream drifted"
    return sd, x


@pytest.fixture(scope="session")
def golden_names():
    from oracle.golden_configs import GOLDEN_CONFIGS
    return list(GOLDEN_CONFIGS)

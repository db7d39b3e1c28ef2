import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


def record_metric(name: str, **values) -> None:
    """Append a measured parity figure to gpurun_out/parity_metrics.jsonl (scratch; copied into profiles/ by hand) so the
    bounds written in the tests can be audited against what was measured (each bound is <= 2x the recorded figure's worst)."""
    import json
    out = ROOT / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        with open(out / "parity_metrics.jsonl", "a") as f:
            f.write(json.dumps({"name": name, **values}) + "\n")
    except OSError:
        pass

"""ORACLE helper: import the unmodified reference -- from /root/reference in the build container, else from the byte-for-byte
copy oracle/build_ref.py shipped as oracle/_ref (the GPU box has no /root/reference).

`src/models/__init__.py` imports `diffusers` and the trainers import `matplotlib`; neither is
installed, so both are stubbed in sys.modules before importing.  Nothing here is copied from the reference;
it is executed in place to produce golden vectors (oracle/make_golden.py) and to validate oracle/*.py.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

_CANDIDATES = (Path("/root/reference"), Path(__file__).resolve().parent / "_ref")


def _root():
    for c in _CANDIDATES:
        if (c / "src" / "models" / "unet.py").exists():
            return c
    return None


REFERENCE_ROOT = _root() or _CANDIDATES[0]


def available() -> bool:
    return _root() is not None


def kind() -> str:
    """'reference' when the unmodified reference can be executed here, else 'port' (callers fall back to oracle/*.py)."""
    return "reference" if available() else "port"


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package
    sys.modules[name] = m
    return m


def install_stubs() -> None:
    class _Missing:  # placeholder classes the reference only names at import time
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party class")

    _stub("diffusers", UNet2DConditionModel=_Missing)
    _stub("diffusers.models")
    _stub("diffusers.models.unets")
    _stub("diffusers.models.unets.unet_2d_condition", UNet2DConditionModel=_Missing)
    _stub("diffusers.models.attention_processor", AttnProcessor2_0=_Missing)
    for name in ("matplotlib", "matplotlib.pyplot"):
        try:
            __import__(name)
        except Exception:
            _stub(name)


def load():
    """Returns a namespace with the reference's UNet, NoiseScheduler (cosine), ImprovedDiffusionTrainer,
    FinalNoiseScheduler (linear) and FinalPokemonGenerator."""
    root = _root()
    if root is None:
        raise RuntimeError("neither /root/reference nor oracle/_ref (python -m oracle.build_ref) is present on this machine")
    install_stubs()
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    from src.models.unet import UNet  # type: ignore
    from src.training.improved_diffusion_trainer import ImprovedDiffusionTrainer, NoiseScheduler  # type: ignore
    from src.training import final_trainer  # type: ignore
    return types.SimpleNamespace(UNet=UNet, NoiseScheduler=NoiseScheduler, ImprovedDiffusionTrainer=ImprovedDiffusionTrainer,
                                 FinalNoiseScheduler=final_trainer.NoiseScheduler,
                                 FinalPokemonGenerator=final_trainer.FinalPokemonGenerator)

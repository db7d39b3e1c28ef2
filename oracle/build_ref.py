"""ORACLE recipe (test / baseline infrastructure, never on the product path): make the UNMODIFIED reference package
available to the GPU box.

    python -m oracle.build_ref          # build container only: needs /root/reference

The reference is pure Python (nothing to compile), but /root/reference does not exist on the GPU box.  This copies its
`src/` package byte for byte into `oracle/_ref/src/` (git-ignored, NOT gpurun-ignored: it travels with the snapshot like a
built .so) together with a manifest of sha256 digests, so that `bench.py --impl reference` / `cpu_baseline` time the
reference's own `ImprovedDiffusionTrainer.train_epoch` and `UNet`, not a port.  No reference source is committed to the
repository's history; oracle/ref_loader.py supplies the import-time stubs for the two absent third-party packages
(diffusers, matplotlib) without touching the copied files.
"""
from __future__ import annotations

import hashlib
import json
import shutil
from pathlib import Path

SRC = Path("/root/reference")
DST = Path(__file__).resolve().parent / "_ref"


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref is present and verified (freshly copied or already matching)."""
    if not (SRC / "src" / "models" / "unet.py").exists():
        return (DST / "MANIFEST.json").exists()
    if DST.exists():
        shutil.rmtree(DST)
    DST.mkdir(parents=True)
    manifest = {}
    for f in sorted((SRC / "src").rglob("*.py")):
        rel = f.relative_to(SRC)
        out = DST / rel
        out.parent.mkdir(parents=True, exist_ok=True)
        data = f.read_bytes()
        out.write_bytes(data)
        manifest[str(rel)] = hashlib.sha256(data).hexdigest()
    (DST / "MANIFEST.json").write_text(json.dumps({"source": str(SRC), "files": manifest}, indent=1))
    if verbose:
        print(f"[oracle] copied {len(manifest)} reference files to {DST}")
    return True


def verify() -> bool:
    """The shipped copy is byte-identical to what the manifest recorded (run on the GPU box before timing it)."""
    mf = DST / "MANIFEST.json"
    if not mf.exists():
        return False
    files = json.loads(mf.read_text())["files"]
    return all((DST / rel).exists() and hashlib.sha256((DST / rel).read_bytes()).hexdigest() == dig for rel, dig in files.items())


if __name__ == "__main__":
    build()
    assert verify()

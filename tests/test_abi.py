"""CPU checks of the C-ABI boundary: the library loads without a GPU/driver and exports every symbol the public
header declares; the ctypes structure mirrors have the C layout; compute calls fail loudly without CUDA."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from pokemon_sprite_generator_b200 import _lib, build
    if not _lib.lib_file().exists():
        build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    text = (ROOT / "include" / "psg_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(psg_[a-z0-9_]+)\s*\(", text)))
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/psg_b200.h but not exported: {missing}"
    assert lib.psg_version() >= 100


def test_ctypes_struct_layout_matches_c():
    from pokemon_sprite_generator_b200 import _lib as L
    # sizes implied by the C declarations on LP64 (natural alignment)
    assert C.sizeof(L.PsgOperand) == 64
    assert C.sizeof(L.PsgEpilogue) == 128
    assert C.sizeof(L.PsgGemmDesc) == 64 * 2 + 24 + 8 + 128
    assert L.PsgEpilogue.alpha.offset == 60 and L.PsgEpilogue.drop_seed.offset == 112


def test_compute_paths_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pokemon_sprite_generator_b200._lib import PsgError
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    from pokemon_sprite_generator_b200.unet import UNet, ResBlock
    with pytest.raises(PsgError):
        NoiseScheduler().add_noise(torch.zeros(1, 8, 27, 27), torch.zeros(1, 8, 27, 27), torch.zeros(1, dtype=torch.long))
    with pytest.raises(PsgError):
        ResBlock(64, 64)(torch.zeros(1))          # containers never compute
    # product code must not import the oracle
    for f in (ROOT / "pokemon_sprite_generator_b200").glob("*.py"):
        assert "oracle" not in f.read_text().replace("the oracle", ""), f


def test_groupnorm_cluster_plan_covers_the_unet_shapes(lib):
    """Host-side planner of the cluster-split GroupNorm kernels (norm_cluster.cu): every GroupNorm shape of the U-Net gets a
    plan whose CTAs tile the unit's pixels exactly, fit in shared memory and use a legal cluster size."""
    out = (C.c_int * 8)()
    for hw, c in [(729, 320), (729, 640), (196, 640), (196, 1280), (49, 1280), (49, 2560), (16, 1280), (16, 2560)]:
        for bwd in (0, 1):
            assert lib.psg_groupnorm_cluster_plan(256, hw, c, 32, bwd, out) == 0, (hw, c, bwd)
            cc, s, rp, r, tu, u, iters, smem = list(out)
            assert cc in (80, 160) and c % cc == 0 and cc % (c // 32) == 0          # whole groups, whole 32 B sectors
            assert s in (1, 2, 4, 8) and (s - 1) * rp < hw <= s * rp                  # every rank owns rows, all rows owned
            assert tu == (cc // 8) * r and tu % 32 == 0 and iters * r >= rp and iters <= 16
            assert (u == 1 or s == 1) and tu * u <= 512 and smem <= 200 * 1024
        assert lib.psg_groupnorm_fused_ok(256, hw, c, 32, 1) == 1
    # shapes outside the family are refused by the cluster planner and served by the slab kernels
    assert lib.psg_groupnorm_cluster_plan(2, 100, 96, 8, 0, out) != 0
    assert lib.psg_groupnorm_fused_ok(2, 100, 96, 8, 1) == 1


def test_groupnorm_stream_plan_and_workspace(lib):
    """Host-side planner of the streaming two-phase GroupNorm backward (norm_stream.cu): chunks tile a sample's pixels exactly,
    thread blocks are whole vector columns x row lanes, sample groups hold the stated bytes, the grid carries both phases of
    every group, and the workspace query covers partial + per-chunk moments + ready counters."""
    lib.psg_groupnorm_bwd_workspace_floats.restype = C.c_longlong
    lib.psg_groupnorm_stream_tune.restype = C.c_longlong
    group_bytes = lib.psg_groupnorm_stream_tune(0, C.c_longlong(-1))
    assert group_bytes > 0
    out = (C.c_int * 8)()
    B = 256
    for hw, c in [(729, 320), (729, 640), (196, 640), (196, 1280), (49, 1280), (49, 2560), (16, 1280), (16, 2560), (100, 96)]:
        g = 32 if c % 32 == 0 and c >= 320 else 8
        assert lib.psg_groupnorm_stream_plan(B, hw, c, g, out) == 0, (hw, c)
        v, rl, t, rows, nchunks, gs, smem, grid = list(out)
        assert v == c // 8 and t == v * rl and 0 < t <= 320 and rl >= 1
        assert (nchunks - 1) * rows < hw <= nchunks * rows
        assert 1 <= gs <= B and (gs == B or gs * hw * c * 4 <= group_bytes < (gs + 1) * hw * c * 4)
        assert grid == -(-B // gs) * 2 * gs * nchunks
        assert smem <= 112 * 1024 and smem >= 3 * rl * c * 4
        need = lib.psg_groupnorm_bwd_workspace_floats(B, hw, c, g)
        assert need >= B * c * 3 + B * nchunks * 3 * c + B
    assert lib.psg_groupnorm_stream_plan(2, 49, 100, 4, out) != 0           # channels not a multiple of 8: refused
    assert lib.psg_groupnorm_bwd_workspace_floats(2, 49, 100, 4) == 2 * 100 * 3


def test_umma_planner_tile_shapes(lib):
    """Host-side tile planner of the tcgen05 engine (psg_umma_plan, no GPU): plain TN products with whole 256-row pair tiles
    take 128-row CTAs (run as cta_group::2 pairs) whatever K is; im2col products with long K keep the 256-row CTA tile; Linear
    weight gradients take 256-row tiles only for long reductions; N picks 256 / 160-wide column tiles."""
    from pokemon_sprite_generator_b200 import _lib as L

    def plan(am, bm, M, N, K, b_conv=False):
        d = L.PsgGemmDesc()
        d.a.mode, d.b.mode, d.M, d.N, d.K, d.in_dtype = am, bm, M, N, K, L.DT_BF16
        bn, mt = C.c_int(0), C.c_int(0)
        assert lib.psg_umma_plan(C.byref(d), C.byref(bn), C.byref(mt)) == 0
        return bn.value, mt.value

    for M, N, K in [(50176, 640, 640), (50176, 1280, 640), (12544, 1280, 2560), (8192, 1280, 2560), (12544, 1280, 3840)]:
        assert plan(L.OP_KMAJOR, L.OP_KMAJOR, M, N, K) == (256, 1), (M, N, K)
    assert plan(L.OP_KMAJOR, L.OP_KMAJOR, 4096, 1280, 2560) == (256, 2)              # 4x4 level: single CTA, long-K rule
    assert plan(L.OP_KMAJOR, L.OP_KMAJOR, 186624, 320, 640) == (160, 1)              # exact 160-wide tiles
    assert plan(L.OP_IM2COL, L.OP_KMAJOR, 50176, 640, 5760) == (256, 2)              # conv fprop / dgrad
    assert plan(L.OP_MNMAJOR, L.OP_MNMAJOR, 1280, 1280, 4096) == (256, 1)            # Linear wgrad, short reduction
    assert plan(L.OP_MNMAJOR, L.OP_MNMAJOR, 1280, 1280, 12544) == (256, 2)
    assert plan(L.OP_MNMAJOR, L.OP_IM2COL_T, 640, 5760, 50176)[1] == 2               # conv wgrad: 256-row tiles despite row padding

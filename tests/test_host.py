"""CPU-side tests: the C-ABI library loads and exports every symbol include/m4d.h declares, the ctypes table covers
them, host logic (new_traj, channel layout, sharding, metric reduction) and the world_size-2 gloo all-gather."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "m4d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(m4d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "m4depth_b200", "libm4d.so"))
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/m4d.h but not exported"
    lib.m4d_abi_version.restype = ctypes.c_int
    assert lib.m4d_abi_version() == 1


def test_ctypes_table_matches_header():
    from m4depth_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_argument_validation_without_gpu():
    """Entry points validate before touching the device: usable on a CPU box, never exit()."""
    from m4depth_b200 import _lib
    assert _lib.lib.m4d_fill(None, 4, 0.0, None) == -1
    assert "null" in _lib.last_error()
    assert _lib.lib.m4d_conv3x3_nhwc(1, 4, 1, 1, 1, 4, 4, 4, 4, 3, 0.1, 1, 4, 0, None) == -1
    assert "stride" in _lib.last_error()
    assert _lib.lib.m4d_sncv_fwd(16, 16, 1, 4, 4, 6, 1, 3, 16, 49, None) == -1
    assert "multiple of 4" in _lib.last_error()


def test_new_entry_points_validate_without_gpu():
    """The stride-aware tensor-core conv calls, BackProjectGrad, the fused first layer and the SNCV variant selector
    reject bad arguments with a return code and a message before any device work."""
    from m4depth_b200 import _lib
    L = _lib.lib
    # packed sizes: [kb][tap][2 planes][cout padded to 16][32 channels] (+32 floats for the weight scale), fp16 = half the floats
    assert L.m4d_conv3x3_tc_packed_floats_p(128, 128, 1, 0) == 4 * 9 * 2 * 128 * 32 + 32
    assert L.m4d_conv3x3_tc_packed_floats_p(128, 128, 1, 1) == 4 * 9 * 2 * 128 * 16 + 32
    assert L.m4d_conv3x3_tc_packed_floats_p(16, 5, 1, 1) == 1 * 9 * 2 * 16 * 16 + 32
    assert L.m4d_conv3x3_tc_packed_floats_p(16, 16, 2, 1) == 2 * 9 * 2 * 16 * 16 + 32          # stride 2: 4C channels in 2 k-blocks
    assert L.m4d_conv3x3_tc_packed_floats_p(192, 192, 2, 1) == 24 * 9 * 2 * 192 * 16 + 32
    assert L.m4d_conv3x3_tc_packed_floats_p(24, 16, 2, 1) == 0                                # stride 2 needs cin % 16 == 0
    assert L.m4d_conv3x3_tc_packed_floats_p(16, 300, 1, 1) == 0 and L.m4d_conv3x3_tc_packed_floats_p(16, 16, 1, 7) == 0
    assert L.m4d_conv3x3_tc_packed_floats(64, 32) == L.m4d_conv3x3_tc_packed_floats_s(64, 32, 1) == L.m4d_conv3x3_tc_packed_floats_p(64, 32, 1, 0)
    assert L.m4d_conv3x3_tc_pack_p(None, 16, 16, 1, 1, None, None) == -1 and "null" in _lib.last_error()
    assert L.m4d_conv3x3_tc_pack_p(16, 24, 16, 2, 1, 16, None) == -1 and "stride" in _lib.last_error()
    assert L.m4d_conv3x3_tc_fwd_p(16, 16, 16, 16, 1, 8, 8, 16, 16, 1, 5, 0.1, 16, 16, 0, None) == -1 and "precision" in _lib.last_error()
    assert L.m4d_conv3x3_tc_fwd_p(16, 16, 16, 16, 1, 7, 8, 16, 16, 2, 1, 0.1, 16, 16, 0, None) == -3     # odd height, stride 2: not supported
    assert "outside the tcgen05 path" in _lib.last_error()
    assert L.m4d_conv3x3_tc_fwd_p(16, 18, 16, 16, 1, 8, 8, 16, 16, 1, 1, 0.1, 16, 16, 0, None) == -3     # pixel stride not a multiple of 4
    dim = (ctypes.c_int32 * 6)(1, 4, 4, 1, 1, 0)
    assert L.m4d_backproject_bwd(16, 16, 16, dim, 16, 16, None) == -1 and "dim[5]" in _lib.last_error()
    assert L.m4d_backproject_bwd(None, 16, 16, dim, 16, 16, None) == -1 and "null" in _lib.last_error()
    assert L.m4d_rgb_conv_dn(16, 2, 16, 16, 1, 8, 8, 16, 16, 0.1, 16, 16, None) == -1 and "bad sizes" in _lib.last_error()
    assert L.m4d_sncv_fwd_ex(16, 16, 1, 4, 4, 8, 1, 2, 16, 25, 0, None) == -1 and "search_range" in _lib.last_error()
    # first encoder layer with the conv weights as kernel parameters: argument checks come before any device work
    assert L.m4d_rgb_conv_dn_hostw(16, 2, 16, 16, 1, 8, 8, 16, 16, 0.1, 16, 16, None) == -1 and "bad sizes" in _lib.last_error()
    assert L.m4d_rgb_conv_dn_hostw(16, 3, None, 16, 1, 8, 8, 16, 16, 0.1, 16, 16, None) == -1 and "null" in _lib.last_error()
    assert L.m4d_rgb_conv_stats_hostw(16, 2, 16, 16, 1, 8, 8, 16, 16, None) == -1 and "bad sizes" in _lib.last_error()
    assert L.m4d_rgb_conv_stats_hostw(16, 3, 16, 16, 1, 8, 8, 24, 16, None) == -1 and "16-byte aligned" in _lib.last_error()
    assert L.m4d_domain_norm_apply(16, 1, 8, 8, 24, 16, 16, 0.1, 16, 16, None) == -1 and "c must be 16 or 32" in _lib.last_error()
    assert L.m4d_domain_norm_apply(16, 1, 8, 8, 16, 16, 16, 0.1, None, 16, None) == -1 and "null" in _lib.last_error()
    assert L.m4d_debug_conv_profile(None) in (0, 1)
    assert L.m4d_pscv_fused_bwd(16, 16, 16, 16, 16, 4, 16, 16, 16, 1, 1, 8, 32, 2, 4, 16, 18, None, 9, 16, 16, 16, 16, None) == -1
    assert "h,w >= 2" in _lib.last_error()
    assert L.m4d_pscv_fused_bwd(16, 16, 16, 16, 16, 4, 16, 16, 16, 1, 8, 8, 32, 2, 4, 16, 17, None, 9, 16, 16, 16, 16, None) == -1
    assert "stride" in _lib.last_error()


def test_no_cpu_fallback():
    import m4depth_b200
    with pytest.raises(m4depth_b200.M4DError):
        m4depth_b200.utils.cost_volume(torch.zeros(1, 4, 4, 8), torch.zeros(1, 4, 4, 8), 3)
    if not torch.cuda.is_available():
        with pytest.raises(m4depth_b200.M4DError):
            m4depth_b200.M4Depth(nbre_levels=3)
    with pytest.raises(NotImplementedError):
        m4depth_b200.M4Depth(nbre_levels=3, is_training=True)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "m4depth_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"


def test_new_traj_flag_and_layout():
    from m4depth_b200.m4depth_network import _new_traj_flag, DepthEstimatorLevel, M4depthAblationParameters
    assert _new_traj_flag([True, False]) and not _new_traj_flag([False, True])
    assert _new_traj_flag(torch.tensor([True])) and not _new_traj_flag(False)
    for depth, cin in ((1, 64), (2, 122), (3, 122), (4, 238), (5, 238), (6, 470)):
        lvl = DepthEstimatorLevel({"is_training": False, "ablation": M4depthAblationParameters()}, depth)
        lvl.build((1, 2, 2, [16, 32, 64, 96, 128, 192][depth - 1]), "cpu")
        assert lvl.cin == cin and lvl.xs % 4 == 0 and lvl.xs >= cin
        assert lvl.ch_logprev == cin - 1 and lvl.ch_sncv == 9 * lvl.nbre_cuts + 5
    lvl = DepthEstimatorLevel({"is_training": False, "ablation": M4depthAblationParameters(SNCV=False, level_memory=False)}, 2)
    lvl.build((1, 2, 2, 32), "cpu")
    assert lvl.cin == 18 + 1 + 1 and lvl.ch_other == -1 and lvl.ch_sncv == -1


def test_shard_sequences():
    from m4depth_b200.dist import shard_sequences
    for n, world in ((64, 8), (7, 2), (3, 4), (0, 2)):
        got = [list(shard_sequences(n, r, world)) for r in range(world)]
        assert sum(got, []) == list(range(n))
        assert max(len(g) for g in got) - min(len(g) for g in got) <= 1


def test_metrics_reduce():
    from m4depth_b200.metrics import MetricsAccumulator, METRIC_NAMES
    p = torch.zeros(2, 14, dtype=torch.float64)
    p[0, :7] = torch.arange(7) * 2.0
    p[0, 7:] = 2
    p[1, :7] = torch.arange(7) * 1.0
    p[1, 7:] = 1
    r = MetricsAccumulator.reduce(p)
    assert list(r) == METRIC_NAMES and abs(r["Delta3"] - 6.0) < 1e-12


WORKER = r"""
import os, sys, torch
sys.path.insert(0, {root!r})
from m4depth_b200 import dist as d
from m4depth_b200.metrics import MetricsAccumulator
rank, world, _ = d.init_process_group("gloo")
seqs = list(d.shard_sequences(5, rank, world))
part = torch.zeros(14, dtype=torch.float64)
part[:7] = float(sum(seqs))
part[7:] = float(len(seqs))
allp = d.all_gather_partials(part)
res = MetricsAccumulator.reduce(allp)
t = d.max_over_ranks(1.0 + rank, "cpu")
assert allp.shape == (world, 14), allp.shape
assert abs(res["AbsRel"] - (0 + 1 + 2 + 3 + 4) / 5.0) < 1e-12, res
assert t == float(world), t
print("rank", rank, "ok")
"""


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2

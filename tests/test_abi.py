"""The C-ABI library: builds, loads, exports every symbol include/muvo_b200.h declares, and validates
arguments on the host (no kernel is launched here -- there is no GPU in the build container)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from muvo_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "muvo_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MUVO_API\s+(?:const\s+char\*|int)\s+(muvo_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = declared_symbols()
    for must in ("muvo_voxelize", "muvo_range_project", "muvo_points_fused", "muvo_bev_pool_fwd", "muvo_bev_pool_bwd",
                 "muvo_segment_sum_fwd", "muvo_segment_sum_bwd", "muvo_ssc_counts", "muvo_ssc_counts_from_logits",
                 "muvo_ws_reset", "muvo_points_workspace_bytes", "muvo_strerror", "muvo_abi_version"):
        assert must in names
    assert names == sorted(_lib.SIGNATURES.keys()), "ctypes table and header out of sync"


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", build.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for name in declared_symbols():
        assert name in exported, name
        assert getattr(lib, name) is not None
    # nothing else leaks out of the library
    assert {n for n in exported if not n.startswith("muvo_")} == set()


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "muvo_b200.h"\nint main(void){ MuvoGrid g; MuvoRangeCfg r; (void)g; (void)r; return MUVO_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_struct_layout_matches_c(tmp_path):
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "muvo_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(MuvoGrid),offsetof(MuvoGrid,size),offsetof(MuvoGrid,roadline_id),sizeof(MuvoRangeCfg),'
                   'offsetof(MuvoRangeCfg,fov),offsetof(MuvoRangeCfg,lidar_pos),sizeof(MuvoLidarPrep),'
                   'offsetof(MuvoLidarPrep,use_ego_box),offsetof(MuvoLidarPrep,remap256));return 0;}\n')
    exe = tmp_path / "s"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.MuvoGrid), _lib.MuvoGrid.size.offset, _lib.MuvoGrid.roadline_id.offset,
            C.sizeof(_lib.MuvoRangeCfg), _lib.MuvoRangeCfg.fov.offset, _lib.MuvoRangeCfg.lidar_pos.offset,
            C.sizeof(_lib.MuvoLidarPrep), _lib.MuvoLidarPrep.use_ego_box.offset, _lib.MuvoLidarPrep.remap256.offset]
    assert got == want


def test_version_and_strerror(lib):
    assert lib.muvo_abi_version() == _lib.ABI_VERSION
    assert lib.muvo_strerror(0) == b"ok"
    assert b"NULL" in lib.muvo_strerror(-1)
    assert b"workspace" in lib.muvo_strerror(-4)


def test_workspace_size_and_argument_errors(lib):
    from muvo_b200.points import GridSpec, RangeSpec
    g, r = GridSpec().to_c(), RangeSpec().to_c()
    n = C.c_size_t(0)
    assert lib.muvo_points_workspace_bytes(100000, 1, C.byref(g), C.byref(r), C.byref(n)) == 0
    # bitmap (G/8) + chunk prefix (G/32) + voxel winner table (8 B/voxel) + pixel table (8 B/px)
    # + queue length / first-frame slots (32 KiB) + rare-path queue (16 B/pt)
    # + dataflow kernel: per-frame claim / completion counters ((6 F + 64) * 4 B) and range-image bin edges ((2 (W + 1) + H + 1) * 8 B)
    G = 192 * 192 * 64
    need = G // 8 + G // 32 + 8 * G + 8 * 64 * 1024 + 32768 + 16 * 100000 + (6 + 64) * 4 + (2 * 1025 + 65) * 8
    assert need <= n.value < need + 4096
    assert lib.muvo_points_workspace_bytes(-1, 1, C.byref(g), C.byref(r), C.byref(n)) == -2
    assert lib.muvo_points_workspace_bytes(10, 1, C.byref(g), C.byref(r), None) == -1
    # NULL / bad-argument paths return before touching the device
    assert lib.muvo_voxelize(None, 0, None, None, 1, 10, None, None, None, None, None, None, None, None, 0, None) == -1
    assert lib.muvo_voxelize(None, 7, None, None, 1, 10, C.byref(g), None, None, None, None, None, None, None, 0, None) == -2
    assert lib.muvo_ssc_counts(None, 0, None, None, None, 0, 10, 2, None, None) == -1
    assert lib.muvo_ssc_counts(None, 0, None, None, None, 0, -5, 2, C.c_void_p(8), None) == -2
    assert lib.muvo_bev_pool_workspace_bytes(1, 100, 0, C.byref(n)) == -2
    assert lib.muvo_segment_sum_workspace_bytes(1000, C.byref(n)) == 0 and n.value > 0


def test_missing_library_is_loud(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MuvoError):
        _lib.load()

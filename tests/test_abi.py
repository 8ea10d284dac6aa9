"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/trackdlo_b200.h declares, its structs match the ctypes mirror, and it fails loudly
(no CPU fallback) when no B200 is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

from trackdlo_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "trackdlo_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tdlo_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in api.ABI_SYMBOLS:
        assert s in syms
    assert set(syms) == set(api.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    for s in _declared_symbols():
        assert hasattr(lib, s), f"{s} missing from {api.LIB_PATH}"
    assert b"sm_100a" in lib.tdlo_version()


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_struct_layout_matches_header():
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "trackdlo_b200.h"
int main(void){
  printf("%zu %zu %zu %zu ", sizeof(tdlo_cpd_params), sizeof(tdlo_track_params), sizeof(tdlo_cpd_batch), sizeof(tdlo_track_batch));
  printf("%zu %zu %zu %zu ", offsetof(tdlo_cpd_params, max_iter), offsetof(tdlo_track_params, max_iter),
         offsetof(tdlo_cpd_batch, status), offsetof(tdlo_track_batch, state));
  printf("%zu %zu %zu %zu ", sizeof(tdlo_vis_batch), offsetof(tdlo_vis_batch, visible_ext_offsets),
         sizeof(tdlo_seq_batch), offsetof(tdlo_seq_batch, status_traj));
  printf("%zu %zu ", sizeof(tdlo_err_batch), offsetof(tdlo_err_batch, error));
  printf("%zu %zu %zu %zu ", sizeof(tdlo_frontend_batch), offsetof(tdlo_frontend_batch, status), offsetof(tdlo_cpd_batch, priors_stride), offsetof(tdlo_track_batch, packed_results));
  printf("%zu %zu %zu ", offsetof(tdlo_vis_batch, proj), offsetof(tdlo_vis_batch, pixel_width), offsetof(tdlo_vis_batch, not_self_occluded));
  printf("%zu %zu\n", offsetof(tdlo_seq_batch, proj), offsetof(tdlo_seq_batch, pixel_width));
  return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c"); exe = os.path.join(d, "t")
        open(src, "w").write(prog)
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        vals = [int(v) for v in subprocess.check_output([exe]).split()]
    assert vals[0] == C.sizeof(api.CpdParamsC) and vals[1] == C.sizeof(api.TrackParamsC)
    assert vals[2] == C.sizeof(api.CpdBatchC) and vals[3] == C.sizeof(api.TrackBatchC)
    assert vals[4] == api.CpdParamsC.max_iter.offset and vals[5] == api.TrackParamsC.max_iter.offset
    assert vals[6] == api.CpdBatchC.status.offset and vals[7] == api.TrackBatchC.state.offset
    assert vals[8] == C.sizeof(api.VisBatchC) and vals[9] == api.VisBatchC.visible_ext_offsets.offset
    assert vals[10] == C.sizeof(api.SeqBatchC) and vals[11] == api.SeqBatchC.status_traj.offset
    assert vals[12] == C.sizeof(api.ErrBatchC) and vals[13] == api.ErrBatchC.error.offset
    assert vals[14] == C.sizeof(api.FrontendBatchC) and vals[15] == api.FrontendBatchC.status.offset
    assert vals[16] == api.CpdBatchC.priors_stride.offset and vals[17] == api.TrackBatchC.packed_results.offset
    assert vals[18] == api.VisBatchC.proj.offset and vals[19] == api.VisBatchC.pixel_width.offset and vals[20] == api.VisBatchC.not_self_occluded.offset
    assert vals[21] == api.SeqBatchC.proj.offset and vals[22] == api.SeqBatchC.pixel_width.offset


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.TdloError) as e:
        api.Context(max_frames=1, max_nodes=30, max_points_total=100)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_create_rejects_bad_capacities():
    lib = api.load_library()
    h = C.c_void_p()
    assert lib.tdlo_create(C.byref(h), 0, 0, 30, 100) == -1
    assert lib.tdlo_create(C.byref(h), 0, 1, 1000, 100) == -1
    assert b"bad capacities" in lib.tdlo_last_error(None)


def test_adapter_header_compiles_against_matrix_stub():
    """The Eigen-facing `class trackdlo` (include/trackdlo_adapter.hpp) must compile with the reference's
    signatures; Eigen is absent here so a minimal column-major MatrixXd stand-in is used."""
    test_src = os.path.join(ROOT, "tests", "cpp", "adapter_compile_test.cpp")
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        obj = os.path.join(d, "a.o")
        subprocess.check_call(["g++", "-std=c++17", "-Wall", "-c", "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "tests", "cpp"), test_src, "-o", obj])

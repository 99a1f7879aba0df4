"""The C-ABI library builds, loads and exports every symbol include/phonic_b200.h declares.
No compute calls (no GPU here); creating a renderer without a device must fail loudly."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "phonic_b200", "csrc", "libphonic_b200.so")
HEADER = os.path.join(ROOT, "include", "phonic_b200.h")


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", os.path.dirname(LIB)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"PB200_API[^;(]*?\b(pb200_\w+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for s in ["pb200_create", "pb200_destroy", "pb200_upload_buffer", "pb200_add_mixer", "pb200_add_effect",
              "pb200_play_file", "pb200_add_sampler", "pb200_schedule", "pb200_render", "pb200_render_device"]:
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_binding_table_matches_header(built_lib):
    from phonic_b200 import _capi
    assert sorted("pb200_" + n for n in _capi.SYMBOLS) == declared_symbols()
    api = _capi.CApi(built_lib, "pb200_")
    assert api.backend().startswith(b"cuda")


def test_oracle_exports_same_surface(oracle_api):
    assert oracle_api.backend().startswith(b"oracle")


def test_struct_sizes_match_c_layout(built_lib):
    # sizeof() of the ctypes mirrors must equal the C compiler's view of the header
    from phonic_b200 import _capi as A
    src = '#include "phonic_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(pb200_config),sizeof(pb200_filter_params),sizeof(pb200_compressor_params),sizeof(pb200_chorus_params),' \
          'sizeof(pb200_reverb_params),sizeof(pb200_file_options),sizeof(pb200_ahdsr),sizeof(pb200_sampler_options),' \
          'sizeof(pb200_event),sizeof(pb200_source_status),sizeof(pb200_voice_state));return 0;}'
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        sizes = list(map(int, subprocess.check_output([os.path.join(d, "t")]).split()))
    mirrors = [A.Config, A.FilterParams, A.CompressorParams, A.ChorusParams, A.ReverbParams, A.FileOptions, A.Ahdsr,
               A.SamplerOptions, A.Event, A.SourceStatus, A.VoiceState]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_create_without_gpu_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from phonic_b200 import _capi as A
    api = A.CApi(built_lib, "pb200_")
    cfg = A.Config(48000, 2, 1024, -1, 1.0)
    r = C.c_void_p()
    assert api.create(C.byref(cfg), C.byref(r)) == A.ERR_CUDA  # no CPU fallback

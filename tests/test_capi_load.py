"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/flan_b200.h declares; without a GPU it refuses loudly instead of falling back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from flan_b200 import build, capi
    build.build_library()
    return capi.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "flan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flan_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from flan_b200 import capi
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(s[0] for s in capi.SYMBOLS) == names


def test_shape_arithmetic_matches_reference(lib):
    assert lib.flan_b200_num_frames(441000, 128) == 3446           # AudioPV.cpp:17
    assert lib.flan_b200_num_frames(28800000, 256) == 112501
    assert lib.flan_b200_analysis_rate(44100.0, 128) == 344.53125  # AudioPV.cpp:25
    assert lib.flan_b200_hop_from_rates(48000.0, 187.5) == 256     # PVBuffer.cpp:381-384


def test_no_cpu_fallback(lib):
    import torch
    from flan_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.FlanB200Error) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)
    from flan_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)


def test_product_does_not_reference_the_oracle():
    # oracle/ is test infrastructure: nothing under flan_b200/ or include/ may import, link or name it
    for base in ("flan_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                    src = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "pv_oracle" not in src and "oracle_lib" not in src and "libflan_ref" not in src, f

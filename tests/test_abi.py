"""CPU suite, part 3: the C-ABI shared library builds for sm_100a, loads without a GPU and exports every symbol that
include/eovae.h declares; the Python binding table lists the same set; the product path refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "eovae.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(eovae_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/eovae.h but not exported"


def test_binding_table_matches_header(built_lib):
    from eo_vae import _C
    assert sorted(_C.SIGNATURES) == _declared_symbols()
    handle = _C.lib()
    assert handle.eovae_version() == 2
    assert handle.eovae_conv_chunk_bytes(128) == 128 and handle.eovae_conv_chunk_bytes(32) == 64
    assert handle.eovae_conv_k_per_tap(12) == 16 and handle.eovae_conv_k_per_tap(512) == 512


def test_sass_is_blackwell_native(built_lib):
    """tcgen05.mma / TMEM loads / TMA show up as UTC*MMA / LDTM / UTMALDG+UTMASTG in the SASS (B200_PROFILING.md)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built_lib], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", "")  # no legacy mma.sync tensor path


def test_no_cpu_fallback():
    from eo_vae import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gn_stats(torch.zeros(1, 32, 4, 4, dtype=torch.bfloat16).to(memory_format=torch.channels_last))
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    with pytest.raises(RuntimeError, match="CUDA"):
        EOConsistencyLoss()(torch.zeros(1, 3, 8, 8), None, torch.zeros(1, 3, 8, 8))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from eo_vae import _C
    monkeypatch.setattr(_C, "_lib", None)
    monkeypatch.setattr(_C, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU"):
        _C.lib()

"""The C-ABI boundary (include/cre.h <-> libcre_b200.so <-> _lib.PROTOTYPES).  No GPU: only loading, symbol
export and the host-only layout / validation entry points are exercised."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from vision_sam3_yolo_lameless_b200 import _lib
from vision_sam3_yolo_lameless_b200.engine import VitConfig

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "cre.h"


def header_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(cre_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_loads():
    assert _lib.LIB_PATH.exists(), "run __graft_entry__.build() first"
    lib = _lib.load()
    assert lib.cre_abi_version() == 2


def test_every_header_symbol_is_exported_and_bound():
    names = header_functions()
    assert len(names) >= 19
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in cre.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with include/cre.h"
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (cre_[a-z0-9_]+)", out))
    assert set(names) <= exported


def test_sass_is_blackwell_native():
    """tcgen05.mma / tcgen05.ld / TMA must be in the shipped SASS (UTC*MMA, LDTM, UTMALDG)."""
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100a".lower() in sass.lower()
    for mnem in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnem in sass, f"{mnem} missing from SASS"
    assert "HMMA.16816" not in sass, "legacy mma.sync path found"


@pytest.mark.parametrize("cfg", [VitConfig.vit_b16(), VitConfig.vit_l16()])
def test_packed_weight_layout(cfg):
    lib = _lib.load()
    cs = cfg.c_struct()
    total = lib.cre_packed_weights_bytes(C.byref(cs))
    d, f = cfg.hidden, cfg.mlp
    mats = d * 768 + cfg.layers * (3 * d * d + d * d + 2 * d * f)
    assert total >= 2 * mats
    spans = []
    for layer, kinds in [(-1, range(0, 5))] + [(l, range(5, _lib.WEIGHT_KINDS)) for l in range(cfg.layers)]:
        for k in kinds:
            off = lib.cre_weight_offset(C.byref(cs), layer, k)
            n = lib.cre_weight_elems(C.byref(cs), layer, k)
            assert off >= 0 and off % 256 == 0 and n > 0
            spans.append((off, off + n * (2 if k in _lib.BF16_KINDS else 4)))
    spans.sort()
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0, "weight tensors overlap"
    assert spans[-1][1] <= total
    assert lib.cre_weight_elems(C.byref(cs), 0, _lib.W_QKV) == 3 * d * d
    assert lib.cre_weight_elems(C.byref(cs), -1, _lib.PREFIX) == 5 * d


def test_argument_validation_without_gpu():
    lib = _lib.load()
    bad = VitConfig(hidden=700).c_struct()
    assert lib.cre_packed_weights_bytes(C.byref(bad)) == -1
    assert "hidden" in _lib.last_error()
    ok = VitConfig.vit_b16().c_struct()
    assert lib.cre_weight_offset(C.byref(ok), 99, _lib.W_QKV) == -1
    assert lib.cre_weight_offset(C.byref(ok), 0, _lib.W_PATCH) == -1       # global kind with a layer index
    assert lib.cre_workspace_bytes(C.byref(ok), 0, 14, 14) == -1
    ws = lib.cre_workspace_bytes(C.byref(ok), 256, 14, 14)
    m = 256 * 201
    assert ws >= m * 768 * 4 + m * 768 * 2 + m * 1536 * 2 + m * 3072 * 2
    assert lib.cre_gallery_scratch_bytes(64, 768, 257) == -1                # k > CRE_TOPK_LIMIT
    assert lib.cre_gallery_scratch_bytes(64, 768, 9) == lib.cre_gallery_scratch_bytes(64, 768, 8) > 0   # k > 8: passes of 8
    assert lib.cre_gallery_scratch_bytes(64, 768, 5) > 0
    assert lib.cre_set_cta_group(3) == -1
    assert lib.cre_destroy(None) == 0
    with pytest.raises(_lib.CreError):
        _lib.check(-1, "probe")


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to compute (and never reaches for the oracle)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from vision_sam3_yolo_lameless_b200.engine import ClipEmbedEngine
    with pytest.raises(_lib.CreError):
        ClipEmbedEngine(VitConfig.vit_b16(), {})
    lib = _lib.load()
    ctx = C.c_void_p()
    cs = VitConfig.vit_b16().c_struct()
    assert lib.cre_create(C.byref(cs), None, 0, C.byref(ctx)) < 0


def test_product_never_imports_oracle():
    pkg = ROOT / "vision_sam3_yolo_lameless_b200"
    for p in pkg.glob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{p.name} imports the oracle"

"""CPU: the C-ABI library loads and exports every symbol include/pixpro_b200.h declares; the
ctypes signature table mirrors the header; argument validation works without a GPU; the torch
wrappers refuse CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pixpro_b200.h")).read()
    return re.findall(r"^PP_API\s+[\w\s\*]+?\b(pp_\w+)\s*\(", src, flags=re.M)


def header_arg_counts():
    src = open(os.path.join(ROOT, "include", "pixpro_b200.h")).read()
    out = {}
    for m in re.finditer(r"^PP_API\s+[\w\s\*]+?\b(pp_\w+)\s*\(([^;]*?)\);", src, flags=re.M | re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args == "void" else len(args.split(","))
    return out


def test_header_declares_the_path():
    syms = header_symbols()
    for need in ["pp_upflow8", "pp_normalize", "pp_concat_flow", "pp_fb_consistency", "pp_flow_stage",
                 "pp_calc_mask_ratio", "pp_add_optical_flow", "pp_regression_loss", "pp_ppm_fwd", "pp_ppm_bwd"]:
        assert need in syms


def test_library_exports_every_declared_symbol():
    from pixpro_b200 import _cabi
    assert os.path.exists(_cabi.LIB_PATH), "libpixpro_b200.so not built: run __graft_entry__.build()"
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for s in header_symbols():
        assert hasattr(raw, s), f"{s} declared in include/pixpro_b200.h but not exported"


def test_ctypes_table_mirrors_header():
    from pixpro_b200 import _cabi
    counts = header_arg_counts()
    assert set(counts) == set(_cabi.SIGNATURES)
    for name, (_, args) in _cabi.SIGNATURES.items():
        assert len(args) == counts[name], f"{name}: header has {counts[name]} args, binding {len(args)}"


def test_abi_version_and_size_queries():
    from pixpro_b200 import _cabi
    L = _cabi.lib()
    assert L.pp_abi_version() == 1
    assert L.pp_regression_loss_workspace(64, 256, 7) == (5 * 64 * 49 + 2 * 64 + 64 * 7) * 4
    assert L.pp_ppm_saved_bytes(2, 256, 49) == (3 * 2 * 49 + 2 * 49 * 49) * 4
    assert L.pp_ppm_bwd_workspace(2, 256, 49) == (3 * 2 * 256 * 49 + 2 * 49 * 49) * 4


def test_argument_validation_needs_no_gpu():
    from pixpro_b200 import _cabi
    L = _cabi.lib()
    assert L.pp_upflow8(None, 1, 4, 4, None, None) == 1  # PP_ERR_INVALID
    assert b"null pointer" in L.pp_last_error()
    with pytest.raises(_cabi.PixProB200Error):
        _cabi.check(L.pp_flow_stage(None, None, 1, 1, 4, 4, 1, 1, 0.01, 0.5, 0, 0, None, None, None, None, None, 0, None), "pp_flow_stage")


def test_wrappers_refuse_cpu_tensors():
    from pixpro_b200 import ops
    from pixpro_b200._cabi import PixProB200Error
    with pytest.raises(PixProB200Error):
        ops.upflow8(torch.zeros(1, 2, 4, 4))
    with pytest.raises(PixProB200Error):
        ops.ppm(torch.zeros(1, 8, 2, 2), torch.zeros(1, 8, 2, 2))
    with pytest.raises(PixProB200Error):
        ops.concat_flow(torch.zeros(2, 1, 2, 4, 4))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pixpro-with-opticalflow_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                for banned in ("import oracle", "from oracle", "libpixpro_oracle", "orc_"):
                    assert banned not in txt, f"{f} references the oracle ({banned})"

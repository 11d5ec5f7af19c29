"""CPU: the C-ABI shared library builds for sm_100a without a GPU, loads, and exports exactly what
include/tqdne_b200.h declares (no compute calls here)."""
import ctypes
import re
import subprocess

import pytest

from tests.conftest import ROOT

HEADER = ROOT / "include" / "tqdne_b200.h"
LIB = ROOT / "tqdne_b200" / "libtqdne_b200.so"


@pytest.fixture(scope="module")
def lib_path():
    if not LIB.exists():
        subprocess.run(["make", "-C", str(ROOT / "tqdne_b200" / "csrc"), "-j8"], check=True, capture_output=True)
    assert LIB.exists()
    return LIB


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(tq_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("tq_plan_add_conv", "tq_plan_add_groupnorm", "tq_plan_add_attention", "tq_plan_add_linear", "tq_plan_run",
                 "tq_plan_add_memset", "tq_edm_euler", "tq_edm_heun", "tq_logspec_griffinlim", "tq_mavg_envelope_inverse",
                 "tq_nchw_to_nhwc", "tq_nhwc_to_nchw", "tq_last_error", "tq_abi_version"):
        assert must in syms


def test_library_exports_every_declared_symbol_and_binding_matches(lib_path):
    from tqdne_b200 import _lib

    handle = ctypes.CDLL(str(lib_path))
    syms = declared_symbols()
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes binding table and header disagree"
    handle.tq_abi_version.restype = ctypes.c_int
    m = re.search(r"#define\s+TQ_ABI_VERSION\s+(\d+)", HEADER.read_text())
    assert handle.tq_abi_version() == int(m.group(1)) == _lib.ABI_VERSION


def test_struct_layouts_match_the_header_field_for_field():
    """ctypes mirrors of the descriptor structs list the header's fields in the header's order."""
    from tqdne_b200 import _lib

    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    bodies = {name: body for body, name in re.findall(r"typedef struct \{([^{}]*)\}\s*(\w+)\s*;", text)}
    for cname, cls in (("tq_conv_desc", _lib.TqConvDesc), ("tq_gn_desc", _lib.TqGnDesc), ("tq_attn_desc", _lib.TqAttnDesc),
                       ("tq_linear_desc", _lib.TqLinearDesc), ("tq_src", _lib.TqSrc), ("tq_slice", _lib.TqSlice),
                       ("tq_gn_bwd_desc", _lib.TqGnBwdDesc)):
        body = bodies[cname]
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?[A-Za-z_0-9]+\s*\*?", "", decl, count=1)  # drop the type
            for part in decl.split(","):
                names.append(re.sub(r"[\*\s]|\[.*\]", "", part))
        assert names == [f[0] for f in cls._fields_], cname


def test_product_path_fails_loudly_without_the_library(monkeypatch, tmp_path):
    from tqdne_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "missing.so")
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.lib()


def test_repack_batch_prepare_is_host_logic(lib_path):
    """tq_repack_batch_prepare only assigns thread-block ranges to the jobs of the one-launch operand repack: callable
    without a GPU.  (Pointers are never dereferenced on the host.)"""
    from tqdne_b200 import _lib

    handle = ctypes.CDLL(str(lib_path))
    handle.tq_repack_batch_prepare.restype = ctypes.c_int64
    handle.tq_repack_batch_prepare.argtypes = [ctypes.POINTER(_lib.TqRepackJob), ctypes.c_int32]
    jobs = (_lib.TqRepackJob * 3)()
    # forward cast of a [128, 5, 192] master: ceil(128 * 5 * 192 / 2048) blocks; two input-gradient copies: one block per
    # 32 x 32 (co, ci) tile and tap
    jobs[0].master, jobs[0].fwd, jobs[0].Op, jobs[0].k, jobs[0].Ip = 0x1000, 0x2000, 128, 5, 192
    jobs[1].master, jobs[1].bwd, jobs[1].Op, jobs[1].k, jobs[1].Ip, jobs[1].ci_off, jobs[1].Cs = 0x1000, 0x3000, 128, 5, 192, 64, 128
    jobs[2].master, jobs[2].bwd, jobs[2].Op, jobs[2].k, jobs[2].Ip, jobs[2].ci_off, jobs[2].Cs = 0x1000, 0x4000, 72, 1, 320, 0, 40
    total = handle.tq_repack_batch_prepare(jobs, 3)
    want = [-(-128 * 5 * 192 // 2048), 4 * 4 * 5, 2 * 3 * 1]
    assert [j.nblocks for j in jobs] == want and total == sum(want)
    assert [j.block0 for j in jobs] == [0, want[0], want[0] + want[1]]
    jobs[1].fwd = 0x5000                       # both copies in one job: refused
    assert handle.tq_repack_batch_prepare(jobs, 3) == -1
    jobs[1].fwd = None
    jobs[2].Cs = 400                           # source slice beyond the master's input channels: refused
    assert handle.tq_repack_batch_prepare(jobs, 3) == -1


def test_attention_writes_lse_decision_is_host_logic(lib_path):
    """Which attention shapes take the multi-block tensor-core kernel -- the one that can leave the rows' log-sum-exp behind
    for the backward pass -- is decided on the host from the descriptor alone."""
    from tqdne_b200 import _lib

    handle = ctypes.CDLL(str(lib_path))
    handle.tq_attention_writes_lse.restype = ctypes.c_int32
    handle.tq_attention_writes_lse.argtypes = [ctypes.POINTER(_lib.TqAttnDesc)]

    def writes(dtype, T, d, causal=0):
        desc = _lib.TqAttnDesc()
        desc.dtype, desc.N, desc.T, desc.heads, desc.d, desc.causal = dtype, 2, T, 4, d, causal
        desc.qkv, desc.out = 0x10000, 0x20000        # 16 B aligned device addresses (never dereferenced here)
        return handle.tq_attention_writes_lse(ctypes.byref(desc))

    BF16, F32 = 0, 1
    assert writes(BF16, 508, 64) == 1 and writes(BF16, 512, 64) == 1 and writes(BF16, 256, 128) == 1   # 1D UNet, pixel UNet
    assert writes(BF16, 300, 64) == 1                  # 384 keys: O fits beside the packed P
    assert writes(BF16, 100, 64) == 0 and writes(BF16, 16, 128) == 0    # one query block / the packed small-T kernel
    assert writes(BF16, 512, 128) == 0                 # K + V of 512 keys x 128 channels do not fit beside two Q buffers
    assert writes(BF16, 600, 64) == 0 and writes(F32, 508, 64) == 0     # FFMA kernels
    assert writes(BF16, 508, 64, causal=1) == 0        # a causal attention runs on the FFMA kernels


"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol include/aga_b200.h declares;
argument validation returns error codes without touching a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def A():
    import __graft_entry__ as G
    G.build()
    import aga_b200
    return aga_b200


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "aga_b200.h")).read()
    return sorted(set(re.findall(r"AGA_API\s+[\w\s\*]+?\b(aga_\w+)\s*\(", txt)))


def test_header_symbols_all_exported_and_bound(A):
    declared = _declared_symbols()
    assert len(declared) >= 17
    lib = A._lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == A._lib.exported_symbols()  # the ctypes table lists exactly the header's symbols


def test_version_and_status_strings(A):
    lib = A._lib.lib()
    assert lib.aga_version() == 100
    assert lib.aga_status_str(0) == b"AGA_OK"
    assert lib.aga_status_str(-2) == b"AGA_ERR_UNSUPPORTED"
    assert lib.aga_status_str(-99) == b"AGA_ERR_UNKNOWN"


def test_argument_validation_without_gpu(A):
    lib = A._lib.lib()
    n = C.c_size_t()
    assert lib.aga_logmel_workspace_bytes(4, 480000, 80, C.byref(n)) == 0 and n.value >= 16
    assert lib.aga_logmel_workspace_bytes(4, 200, 80, C.byref(n)) == -1  # needs N > 200
    assert lib.aga_logmel_workspace_bytes(0, 480000, 80, C.byref(n)) == -1
    assert lib.aga_logmel_packed_filter_bytes(80, C.byref(n)) == 0 and n.value == 16 + 80 * (12 + 804)
    assert lib.aga_logmel_packed_filter_bytes(1000, C.byref(n)) == -1
    assert lib.aga_guided_loss_workspace_bytes(12, 16, 12, C.byref(n)) == 0 and n.value >= 12 * 16 * 12 * 4
    p = A._lib.AttnParams()
    assert lib.aga_attn_fwd_workspace_bytes(C.byref(p), C.byref(n)) == -1  # null pointers
    p.q = p.k = p.v = p.out = 256
    p.B, p.H, p.Tq, p.Tk = 1, 2, 8, 9
    p.causal = 1
    assert lib.aga_attn_fwd_workspace_bytes(C.byref(p), C.byref(n)) == -1  # causal needs Tq == Tk
    p.causal = 0
    p.q_stride_b, p.q_stride_t = 8 * 128, 128
    p.k_stride_b = p.v_stride_b = 9 * 128
    p.k_stride_t = p.v_stride_t = p.o_stride_t = 128
    p.o_stride_b = 8 * 128
    assert lib.aga_attn_fwd_workspace_bytes(C.byref(p), C.byref(n)) == 0
    p.q_stride_t = 130  # rows not 16-byte aligned
    assert lib.aga_attn_fwd_workspace_bytes(C.byref(p), C.byref(n)) == -2


def test_cpu_tensors_are_rejected_loudly(A):
    import torch
    with pytest.raises(A.AgaError):
        A.log_mel_spectrogram(torch.zeros(1, 16000))
    with pytest.raises(A.AgaError):
        A.qkv_attention(torch.zeros(1, 4, 64), torch.zeros(1, 4, 64), torch.zeros(1, 4, 64), 1)


def test_mel_filterbank_generator_matches_reference_npz(A, golden_dir):
    import numpy as np
    ref = np.load(os.path.join(golden_dir, "mel_80_ref.npy"))
    mine = A.mel_filterbank_numpy(80)
    assert np.array_equal(mine != 0, ref != 0) and np.abs(mine - ref).max() < 5e-9

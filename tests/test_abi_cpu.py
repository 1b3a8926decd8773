"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol
include/nkbk.h declares, and its host-only coefficient helpers agree with the
oracle.  No compute entry point is called here (no GPU)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import preprocess as opre

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "nkbk.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nkbk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(nkbk_lib):
    from nkb_classification_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(nkbk_lib, n), f"{n} declared in nkbk.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes signature in _lib.SYMBOLS"
    assert set(_lib.SYMBOLS) == set(names)
    assert nkbk_lib.nkbk_abi_version() == 1


def test_library_is_sm100a_only(nkbk_lib):
    import subprocess
    from nkb_classification_b200._lib import LIB_PATH
    out = subprocess.run(["cuobjdump", "-lelf", str(LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("dsize,ssize", [(224, 264), (224, 48), (224, 480), (224, 448), (224, 224), (128, 5),
                                         (33, 1), (224, 1080), (224, 1920), (17, 399), (256, 255), (7, 2)])
def test_axis_tables_match_oracle(nkbk_lib, dsize, ssize):
    from nkb_classification_b200 import ops
    for horizontal in (True, False):
        s, c0, c1 = ops.debug_axis_table(dsize, ssize, horizontal)
        es, e0, e1 = opre.axis_tables(dsize, ssize, horizontal)
        assert np.array_equal(s, es) and np.array_equal(c0, e0) and np.array_equal(c1, e1)


def test_axis_tables_match_oracle_sweep(nkbk_lib):
    from nkb_classification_b200 import ops
    rng = np.random.default_rng(5)
    for _ in range(300):
        d, s_ = int(rng.integers(1, 300)), int(rng.integers(1, 2000))
        for horizontal in (True, False):
            got = ops.debug_axis_table(d, s_, horizontal)
            exp = opre.axis_tables(d, s_, horizontal)
            assert all(np.array_equal(a, b) for a, b in zip(got, exp)), (d, s_, horizontal)


def test_letterbox_geometry_matches_oracle(nkbk_lib):
    from nkb_classification_b200 import ops
    rng = np.random.default_rng(6)
    cases = [(95, 1, 40, 40, 40)]
    for _ in range(500):
        h, w = int(rng.integers(1, 1100)), int(rng.integers(1, 1950))
        S = int(rng.choice([40, 128, 224, 256]))
        cases.append((h, w, S, S, S))
    cases += [(250, 125, 224, 224, 224), (3, 2, 5, 5, 5), (5, 3, 224, 224, 256), (7, 7, 7, 7, 7)]
    for h, w, ms, oh, ow in cases:
        exp = opre.letterbox_geometry(h, w, ms, oh, ow)
        if exp[0] < 1 or exp[1] < 1:
            with pytest.raises(NotImplementedError):
                ops.debug_letterbox(h, w, ms, oh, ow)
        else:
            assert ops.debug_letterbox(h, w, ms, oh, ow) == exp, (h, w, ms, oh, ow)


def test_argument_errors_need_no_gpu(nkbk_lib):
    """Argument validation happens before any CUDA call and maps to the reference's exception types."""
    from nkb_classification_b200 import _lib
    seg = (ctypes.c_int32 * 2)(0, 3)
    rc = nkbk_lib.nkbk_argmax_confusion(None, 7, 4, 3, seg, 1, None, None, None, None)
    assert rc == _lib.NKBK_E_ARG
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert b"dtype" in nkbk_lib.nkbk_last_error()
    assert nkbk_lib.nkbk_comm_world() == 0
    rc = nkbk_lib.nkbk_allreduce_heads(None, 0, None, 0, None)
    assert rc == _lib.NKBK_E_NCCL
    # K4' (peer memory): not connected -> refused before any CUDA call; bad world -> argument error
    assert nkbk_lib.nkbk_peer_world() == 0
    rc = nkbk_lib.nkbk_peer_allreduce_finalize(None, 8, seg, 1, None, None, None, 0, None)
    assert rc == _lib.NKBK_E_NCCL and b"not connected" in nkbk_lib.nkbk_last_error()
    hbuf = (ctypes.c_uint8 * _lib.IPC_HANDLE_BYTES)()
    assert nkbk_lib.nkbk_peer_init(0, 99, 0, 16, 0, hbuf) == _lib.NKBK_E_ARG
    assert nkbk_lib.nkbk_peer_init(2, 2, 0, 16, 0, hbuf) == _lib.NKBK_E_ARG
    assert nkbk_lib.nkbk_peer_connect(None) == _lib.NKBK_E_ARG
    st = ctypes.c_int32(7)
    assert nkbk_lib.nkbk_peer_status(ctypes.byref(st)) == 0 and st.value == 0
    assert nkbk_lib.nkbk_peer_disconnect() == 0 and nkbk_lib.nkbk_peer_shutdown() == 0


def test_ops_refuse_cpu_tensors(nkbk_lib):
    import torch
    from nkb_classification_b200 import ops, transforms as T
    plan = T.compile_pipeline([T.Resize(8, 8), T.Normalize(), T.ToTensorV2()])
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.preprocess_crops(torch.zeros(1, 4, 4, 3, dtype=torch.uint8), torch.zeros(1, 4, dtype=torch.int32),
                             torch.zeros(1, dtype=torch.int32), plan)

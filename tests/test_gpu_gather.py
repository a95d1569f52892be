"""The row-gather kernel of the partial-mask path (C-ABI actk_gather_rows) against torch.index_select — the reference's
`xz[:, idx, :]` (src/models/base/mamba_layer.py:1963, 1974).  Bit-exact: it is a copy."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(3, 324, 2560, 30), (2, 5184, 640, 1551), (5, 64, 128, 64), (1, 1296, 96, 1), (4, 100, 8, 37)])
def test_gather_rows_equals_index_select(shape, dtype):
    from actalker_b200 import mamba_layer as ml
    B, rows, width, n = shape
    g = torch.Generator(device="cuda").manual_seed(rows + n)
    t = torch.randn(B, rows, width, device="cuda", generator=g).to(dtype)
    idx = torch.randperm(rows, device="cuda", generator=g)[:n].sort().values.int()
    got = ml._gather_rows(t, idx)
    assert got.shape == (B, n, width) and torch.equal(got, t.index_select(1, idx.long()))


def test_gather_rows_rejects_bad_arguments_loudly():
    from actalker_b200 import _lib
    lib = _lib.load()
    t = torch.zeros(2, 8, 8, device="cuda", dtype=torch.bfloat16)
    idx = torch.zeros(4, device="cuda", dtype=torch.int32)
    with pytest.raises(RuntimeError, match="NULL"):
        _lib.check(lib.actk_gather_rows(None, idx.data_ptr(), t.data_ptr(), 2, 8, 4, 16, None), "actk_gather_rows")
    with pytest.raises(RuntimeError, match="multiples of 16"):
        _lib.check(lib.actk_gather_rows(t.data_ptr(), idx.data_ptr(), t.data_ptr(), 2, 8, 4, 8, None), "actk_gather_rows")

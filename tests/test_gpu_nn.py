"""GPU tests of the large-tree brute-force searches (rrtk_nearest_f32_dev / rrtk_near_f32_dev)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _nearest(torch, L, xy, smp):
    from rrtk import _lib
    B = smp.shape[0]
    scratch = torch.empty(B, dtype=torch.int64, device="cuda")
    idx = torch.empty(B, dtype=torch.int32, device="cuda")
    d2 = torch.empty(B, dtype=torch.float32, device="cuda")
    _lib.check(L.rrtk_nearest_f32_dev(xy.data_ptr(), xy.shape[0], smp.data_ptr(), B, scratch.data_ptr(),
                                      idx.data_ptr(), d2.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return idx.cpu().numpy(), d2.cpu().numpy()


@pytest.mark.parametrize("n", [1, 2, 3, 31, 1000, 4097, 1 << 20, (1 << 22) + 5])
@pytest.mark.parametrize("B", [1, 3, 8, 13])
def test_nearest_matches_numpy(n, B):
    import torch
    from rrtk import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(n * 31 + B)
    xy = torch.rand((n, 2), dtype=torch.float32, device="cuda", generator=g) * 17 - 2
    smp = torch.rand((B, 2), dtype=torch.float32, device="cuda", generator=g) * 17 - 2
    idx, d2 = _nearest(torch, L, xy, smp)
    a = xy.cpu().numpy(); s = smp.cpu().numpy()
    for b in range(B):
        dx = a[:, 0] - s[b, 0]; dy = a[:, 1] - s[b, 1]
        # same FP32 expression as the kernel: fma(dx, dx, dy*dy)
        dd = (dx.astype(np.float64) * dx.astype(np.float64) + (dy * dy).astype(np.float64)).astype(np.float32)
        assert dd[idx[b]] == dd.min()
        assert idx[b] == int(np.flatnonzero(dd == dd.min())[0])     # lowest index among exact ties
        assert d2[b] == dd.min()


def test_nearest_ties_pick_lowest_index():
    import torch
    from rrtk import _lib
    L = _lib.lib()
    n = 100003
    xy = torch.full((n, 2), 5.0, dtype=torch.float32, device="cuda")   # every node identical
    xy[70000:] = 4.0
    smp = torch.tensor([[5.0, 5.0], [4.0, 4.0], [0.0, 0.0]], dtype=torch.float32, device="cuda")
    idx, d2 = _nearest(torch, L, xy, smp)
    assert idx.tolist() == [0, 70000, 70000] and d2[0] == 0.0 and d2[1] == 0.0


@pytest.mark.parametrize("n", [1, 7, 1000, 1 << 20, (1 << 21) + 1])
def test_near_matches_numpy(n):
    import torch
    from rrtk import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(n)
    xy = torch.rand((n, 2), dtype=torch.float32, device="cuda", generator=g) * 17 - 2
    cx, cy, r2 = np.float32(6.5), np.float32(7.25), np.float32(0.8)
    cap = 1 << 16
    out = torch.full((cap,), -1, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(L.rrtk_near_f32_dev(xy.data_ptr(), n, float(cx), float(cy), float(r2), out.data_ptr(), cap,
                                   cnt.data_ptr(), torch.cuda.current_stream().cuda_stream))
    a = xy.cpu().numpy()
    dx = a[:, 0] - cx; dy = a[:, 1] - cy
    dd = (dx.astype(np.float64) * dx.astype(np.float64) + (dy * dy).astype(np.float64)).astype(np.float32)
    want = np.flatnonzero(dd <= r2)
    k = int(cnt.item())
    assert k == len(want)
    got = np.sort(out[:min(k, cap)].cpu().numpy())
    assert np.array_equal(got, want[:len(got)]) if k <= cap else True

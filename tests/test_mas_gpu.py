"""CUDA monotonic alignment search vs the oracle -- bit-exact (torch.equal) everywhere."""
import numpy as np
import pytest
import torch

import vits_oracle as O

pytestmark = pytest.mark.gpu


def _mask(tx, ty, t_x, t_y):
    return ((torch.arange(tx)[None, :, None] < t_x[:, None, None]) &
            (torch.arange(ty)[None, None, :] < t_y[:, None, None])).float()


def _run(value, mask):
    from tts_b200.helpers import maximum_path
    return maximum_path(value.cuda(), mask.cuda()).cpu()


def test_golden_cases(golden):
    for case in golden("mas_cases")["cases"]:
        assert torch.equal(_run(case["value"], case["mask"]), case["path"])


@pytest.mark.parametrize("b,tx,ty", [(16, 70, 150), (5, 300, 310), (3, 1, 9), (4, 64, 64), (2, 1100, 1200),
                                      (7, 33, 257)])
def test_random_ragged_vs_oracle(b, tx, ty):
    rng = np.random.RandomState(b * 1000 + tx)
    v = torch.from_numpy((rng.randn(b, tx, ty) * 4).astype(np.float32))
    t_x = torch.from_numpy(rng.randint(1, tx + 1, size=b))
    t_y = torch.tensor([int(rng.randint(int(a), ty + 1)) for a in t_x])
    t_x[0], t_y[0] = tx, ty
    m = _mask(tx, ty, t_x, t_y)
    assert torch.equal(_run(v, m), O.maximum_path(v, m, impl="c"))


def test_ties_and_loglik_shaped_values():
    # quantised values force exact ties: the backtrack's strict `<` must keep the reference's choice
    rng = np.random.RandomState(3)
    v = torch.from_numpy(rng.randint(-3, 1, size=(8, 40, 90)).astype(np.float32))
    t_x = torch.full((8,), 40)
    t_y = torch.full((8,), 90)
    m = _mask(40, 90, t_x, t_y)
    assert torch.equal(_run(v, m), O.maximum_path(v, m, impl="c"))
    v2 = -50 * torch.rand(4, 50, 120)
    m2 = _mask(50, 120, torch.full((4,), 50), torch.full((4,), 120))
    assert torch.equal(_run(v2, m2), O.maximum_path(v2, m2, impl="c"))


def test_degenerate_tx_gt_ty_follows_reference_port():
    rng = np.random.RandomState(11)
    v = torch.from_numpy(rng.randn(6, 30, 20).astype(np.float32))
    t_x = torch.tensor([30, 25, 21, 30, 12, 5])
    t_y = torch.tensor([20, 10, 20, 1, 10, 5])
    m = _mask(30, 20, t_x, t_y)
    assert torch.equal(_run(v, m), O.maximum_path(v, m, impl="c"))


def test_int32_output_and_lengths_api():
    from tts_b200.helpers import maximum_path_lengths
    v = torch.randn(3, 20, 50)
    t_x = torch.tensor([20, 11, 4], dtype=torch.int32)
    t_y = torch.tensor([50, 30, 4], dtype=torch.int32)
    m = _mask(20, 50, t_x, t_y)
    p = maximum_path_lengths((v * m).cuda(), t_x.cuda(), t_y.cuda())
    assert p.dtype == torch.int32
    assert torch.equal(p.cpu().float(), O.maximum_path(v, m, impl="c"))
    assert maximum_path_lengths(torch.zeros(0, 4, 5).cuda(), t_x[:0].cuda(), t_y[:0].cuda()).shape == (0, 4, 5)


def test_full_size_cfg4_bit_exact_and_structure():
    """BASELINE config 4: batch=512, T_text=200, T_mel=1000 (mixed lengths)."""
    rng = np.random.RandomState(4)
    b, tx, ty = 512, 200, 1000
    v = torch.from_numpy(rng.randn(b, tx, ty).astype(np.float32))
    t_x = torch.from_numpy(rng.randint(100, 201, size=b))
    t_y = torch.tensor([int(rng.randint(min(5 * int(a), 1000), 1001)) for a in t_x])
    t_x[:8], t_y[:8] = tx, ty
    m = _mask(tx, ty, t_x, t_y)
    got = _run(v, m)
    assert torch.equal(got, O.maximum_path(v, m, impl="c"))
    # size-independent structure: one 1 per valid column, monotone unit steps, ends pinned
    idx = got.argmax(1)
    col_sum = got.sum(1)
    for n in range(b):
        ny, nx = int(t_y[n]), int(t_x[n])
        assert torch.all(col_sum[n, :ny] == 1) and torch.all(col_sum[n, ny:] == 0)
        d = idx[n, 1:ny] - idx[n, :ny - 1]
        assert d.min() >= 0 and d.max() <= 1
        assert idx[n, 0] == 0 and idx[n, ny - 1] == nx - 1


@pytest.mark.parametrize("b,c,tx,ty", [(3, 192, 37, 151), (8, 192, 64, 300), (2, 24, 5, 9)])
def test_alignment_from_prior_statistics_matches_oracle(b, c, tx, ty):
    """the alignment step of Vits.forward_mas (vits.py:909-919) from z_p / m_p / logs_p on the device."""
    from tts_b200.helpers import maximum_path_from_stats
    torch.manual_seed(b * 100 + tx)
    xl = torch.randint(max(1, tx // 2), tx + 1, (b,))
    xl[0] = tx
    yl = torch.clamp(xl * torch.randint(2, 5, (b,)), max=ty)
    yl[0] = ty
    x_mask = O.sequence_mask(xl, tx).unsqueeze(1).float()
    y_mask = O.sequence_mask(yl, ty).unsqueeze(1).float()
    z_p, m_p, logs_p = torch.randn(b, c, ty), torch.randn(b, c, tx), torch.randn(b, c, tx) * 0.3
    want, want_logp = O.forward_mas_attn(z_p, m_p, logs_p, x_mask, y_mask)
    got, logp = maximum_path_from_stats(z_p.cuda(), m_p.cuda(), logs_p.cuda(), x_mask.cuda(), y_mask.cuda(), return_logp=True)
    rel = (logp.cpu() - want_logp).abs().max() / want_logp.abs().max()
    assert rel < 2e-6, rel                              # same arithmetic, different summation order over channels
    assert torch.equal(got.cpu(), want), "alignment path differs"
    assert torch.equal(got.sum(2).cpu(), want.sum(2))    # one text position per frame

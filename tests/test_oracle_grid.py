"""Hand-computed known answers for the grid restatement (oracle/grid_oracle.py
follows gridencoder.cu:50-84,137-197; the reference ships no vectors for it)."""
import numpy as np
import torch

from oracle import grid_oracle as go


def test_level_sizes_match_survey():
    offs, res = go.make_offsets(3, 10, 2.0, 16, 21)
    assert list(res[:3]) == [17, 33, 65] and res[-1] == 8193
    assert np.diff(offs).tolist()[:3] == [4920, 35944, 274632]
    assert all(n == 2 ** 21 for n in np.diff(offs)[3:])
    assert offs[-1] == 14995560


def test_dense_and_hashed_indices():
    offs, _ = go.make_offsets(3, 10, 2.0, 16, 21)
    # level 0: scale 15, resolution 16, stride 17 -> dense x + 17 y + 289 z
    pg = torch.tensor([[3, 5, 7]], dtype=torch.int64)
    assert int(go.grid_index(pg, 4920, 16)) == 3 + 17 * 5 + 289 * 7
    # level 9: scale 8191, resolution 8192 -> hashed, uint32 wrap
    x, y, z = 8000, 123, 4567
    want = (x ^ ((y * 2654435761) & 0xFFFFFFFF) ^ ((z * 805459861) & 0xFFFFFFFF)) % (2 ** 21)
    assert int(go.grid_index(torch.tensor([[x, y, z]]), 2 ** 21, 8192)) == want


def test_interpolation_of_a_linear_field_is_exact():
    """A table holding f(v) = a.v + b on a dense level is reproduced exactly by
    trilinear interpolation (align_corners=False offsets by 0.5 cell)."""
    offs, res = go.make_offsets(3, 1, 2.0, 16, 21)
    R = 17
    idx = np.arange(R ** 3)
    v = np.stack([idx % R, (idx // R) % R, idx // (R * R)], -1).astype(np.float32)
    table = np.zeros((int(offs[-1]), 1), np.float32)
    table[:R ** 3, 0] = v @ np.array([0.5, -1.0, 2.0], np.float32) + 3.0
    x = torch.rand(256, 3)
    out, _ = go.grid_encode_forward(x, torch.from_numpy(table), torch.from_numpy(offs), 1.0, 16)
    pos = x.double().numpy() * 15 + 0.5
    want = pos @ np.array([0.5, -1.0, 2.0]) + 3.0
    assert np.allclose(out[0, :, 0].numpy(), want, atol=1e-4)


def test_out_of_range_points_give_zero_and_backward_matches_autograd():
    offs, _ = go.make_offsets(3, 4, 2.0, 16, 12)
    offs_t = torch.from_numpy(offs)
    emb = torch.randn(int(offs[-1]), 2)
    x = torch.rand(64, 3)
    x[0, 0] = 1.5
    x[1, 2] = -0.1
    out, _ = go.grid_encode_forward(x, emb, offs_t, 1.0, 16)
    assert torch.all(out[:, :2] == 0)
    g = torch.randn_like(out)
    ge, _ = go.grid_encode_backward(g, x, emb, offs_t, 1.0, 16)
    # forward is linear in the table: <g, F(e)> == <ge, e>
    assert abs(float((g * out).sum()) - float((ge * emb).sum())) < 1e-2 * float((g * out).abs().sum())

import numpy as np


def channels(img):
    img = np.asarray(img, dtype=np.uint32)
    return np.stack([(img >> 16) & 255, (img >> 8) & 255, img & 255, img >> 24], -1).astype(np.int32)


def diff_stats(a, b):
    """(#pixels that differ, max per-channel |delta|, #pixels differing by more than 1 LSB)."""
    ca, cb = channels(a), channels(b)
    d = np.abs(ca - cb).max(-1)
    return int((d > 0).sum()), int(d.max()) if d.size else 0, int((d > 1).sum())


def assert_parity(got, want, what, exact=True):
    """North-star bar: within +-1 LSB per RGB channel. `exact` additionally demands bit-equality
    (what the integer/byte modes must meet, and what the float modes achieve in practice)."""
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    n, mx, over = diff_stats(got, want)
    assert over == 0 and mx <= 1, f"{what}: {n} pixels differ, max channel delta {mx}, {over} pixels beyond 1 LSB"
    if exact:
        assert n == 0, f"{what}: {n} pixels differ by 1 LSB (bit-exactness expected)"

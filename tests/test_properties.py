"""Property tests (hypothesis) of the host-side index and metric logic: invariants that hold for every size, complementing
the tabulated reference cases in tests/golden/."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from dynavsr_b200 import dist as D
from dynavsr_b200.clips import PADDINGS, index_generation
from dynavsr_b200.driver import psnr_from_sse, ssim_u8


@settings(max_examples=300, deadline=None)
@given(st.sampled_from(PADDINGS), st.sampled_from([1, 3, 5, 7]), st.integers(2, 60), st.data())
def test_index_generation_invariants(padding, N, max_n, data):
    if max_n < N:
        max_n = N
    crt = data.draw(st.integers(0, max_n - 1))
    idx = index_generation(crt, max_n, N, padding)
    assert len(idx) == N and idx[N // 2] == crt                       # the centre frame is always the requested one
    assert all(0 <= i < max_n for i in idx)                           # every padding rule stays inside the clip
    half = N // 2
    if half <= crt < max_n - half:                                    # interior windows are plain consecutive frames
        assert idx == list(range(crt - half, crt + half + 1))
    inside = [i for k, i in enumerate(idx) if 0 <= crt - half + k < max_n]
    assert inside == [crt - half + k for k in range(N) if 0 <= crt - half + k < max_n]
    if padding == 'new_info' and max_n >= 2 * N:
        assert len(set(idx)) == N                                     # 'new_info' never repeats a frame


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 500), st.integers(1, 16))
def test_shard_indices_partition_every_item_exactly_once(n, world):
    shards = [D.shard_indices(n, r, world) for r in range(world)]
    assert sorted(i for s in shards for i in s) == list(range(n))
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 10 ** 9), st.integers(1, 10 ** 9), st.integers(1, 10 ** 7))
def test_psnr_from_sse_is_monotone_and_matches_definition(a, b, n):
    pa, pb = psnr_from_sse(a, n), psnr_from_sse(b, n)
    assert (pa > pb) == (a < b) or a == b or math.isclose(pa, pb)
    assert math.isclose(pa, 10 * math.log10(255.0 ** 2 / (a / n)), rel_tol=1e-12, abs_tol=1e-9)


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(11, 24), st.integers(11, 24))
def test_ssim_bounds_symmetry_identity(seed, h, w):
    rng = np.random.RandomState(seed)
    a = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    b = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    s = ssim_u8(a, b)
    assert -1.0 <= s <= 1.0 and math.isclose(s, ssim_u8(b, a), rel_tol=1e-12, abs_tol=1e-12)
    assert math.isclose(ssim_u8(a, a), 1.0, rel_tol=1e-12)

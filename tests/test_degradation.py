"""HR -> LR degradation: oracle vs the golden produced by the unmodified reference class (CPU), and the
CUDA kernel vs both (GPU)."""
import numpy as np
import pytest
import torch

from util import gold, rel

CASES = ['iso', 'aniso', 'x2', 'delta']


@pytest.mark.parametrize('tag', CASES)
def test_oracle_matches_reference_golden(tag):
    from oracle import degradation_oracle as DO
    g = gold('degradation.npz')
    ks, sc, s0, s1, theta = g[tag + '_cfg']
    k = DO.gaussian_kernel(int(ks), [s0, s1], theta)
    assert np.abs(k - g[tag + '_kernel']).max() == 0.0
    y = DO.degrade(torch.from_numpy(g[tag + '_img']), k, int(sc))
    assert torch.equal(y, torch.from_numpy(g[tag + '_out']))


@pytest.mark.parametrize('tag', ['perframe3', 'perframe5'])
def test_oracle_per_frame_kernels(tag):
    from oracle import degradation_oracle as DO
    g = gold('degradation.npz')
    y = DO.degrade(torch.from_numpy(g[tag + '_img']), g[tag + '_kernel'], 4)
    assert torch.equal(y, torch.from_numpy(g[tag + '_out']))


def test_host_kernel_construction_matches_reference_golden():
    from dynavsr_b200.degradation import Degradation
    from oracle import degradation_oracle as DO
    g = gold('degradation.npz')
    for tag in CASES:
        ks, sc, s0, s1, theta = g[tag + '_cfg']
        d = Degradation(int(ks), int(sc), theta=theta, sigma=[s0, s1])
        assert np.abs(d.get_kernel() - g[tag + '_kernel']).max() == 0.0
        assert np.abs(d.kernel_shift(d.get_kernel()) - DO.shift_kernel(g[tag + '_kernel'], int(sc))).max() == 0.0
    with pytest.raises(NotImplementedError):
        Degradation(21, 4).apply(torch.rand(1, 3, 32, 32))


@pytest.mark.gpu
@pytest.mark.parametrize('tag', CASES + ['perframe3', 'perframe5'])
def test_cuda_degradation_matches_reference_golden(tag):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dynavsr_b200.degradation import Degradation
    g = gold('degradation.npz')
    img = torch.from_numpy(g[tag + '_img'])
    if tag.startswith('perframe'):
        d = Degradation(21, 4)
        d.set_kernel_directly(g[tag + '_kernel'])
    else:
        ks, sc, s0, s1, theta = g[tag + '_cfg']
        d = Degradation(int(ks), int(sc), theta=theta, sigma=[s0, s1])
    ref = torch.from_numpy(g[tag + '_out'])
    y = d.apply(img.cuda())
    assert y.shape == ref.shape and rel(y, ref) < 1e-6
    assert rel(d.apply(img[0].cuda()), ref[0]) < 1e-6 if not tag.startswith('perframe') else True      # [C, H, W] form
    # quantised channels-last form used by the synthetic pipeline: at most one 8-bit step away, almost everywhere equal
    yq = d.apply_nhwc(img.cuda().permute(0, 2, 3, 1).contiguous(), quantize=True).permute(0, 3, 1, 2).cpu()
    rq = (ref * 255).round() / 255
    diff = (yq - rq).abs()
    assert float(diff.max()) <= 1.0 / 255 + 1e-6 and float((diff > 1e-6).float().mean()) < 1e-3


@pytest.mark.gpu
def test_cuda_degradation_full_size_properties():
    """REDS shape 5 x 3 x 720 x 1280 -> 180 x 320: constant images stay constant (kernel sums to 1), linearity."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dynavsr_b200.degradation import Degradation
    d = Degradation(21, 4, sigma=[1.6, 1.6])
    a = torch.rand(5, 720, 1280, 3, device='cuda')
    b = torch.rand(5, 720, 1280, 3, device='cuda')
    ya, yb = d.apply_nhwc(a), d.apply_nhwc(b)
    assert ya.shape == (5, 180, 320, 3)
    assert float((d.apply_nhwc(torch.full_like(a, 0.37)) - 0.37).abs().max()) < 1e-5
    assert rel(d.apply_nhwc(2.0 * a - 0.5 * b), 2.0 * ya - 0.5 * yb) < 1e-5


def test_per_frame_kernels_of_different_shifted_size_are_padded_to_a_common_one():
    """kernel_shift pads by an amount that depends on the kernel's own shift, so per-frame kernels can come back with
    different sizes (the reference applies them one at a time, random_kernel_generator.py:100-118); they are zero-padded
    symmetrically to the largest one, which leaves every centre -- and therefore every filtered frame -- unchanged."""
    from scipy import ndimage
    from dynavsr_b200.degradation import Degradation
    d = Degradation(21, 4)
    rng = np.random.RandomState(0)
    ks = []
    for _ in range(3):
        k = np.zeros((21, 21))
        k[10 + rng.randint(-6, 7), 10 + rng.randint(-6, 7)] = 1.0
        k = ndimage.gaussian_filter(k, 1.5)
        ks.append(k / k.sum())
    own = [d.kernel_shift(k) for k in ks]
    assert len({k.shape for k in own}) > 1                       # the case the plain np.stack could not handle
    d.set_kernel_directly(np.stack(ks))
    dev, L = d._device_kernels(torch.device('cpu'))
    assert tuple(dev.shape) == (3, L, L) and L == max(k.shape[0] for k in own)
    for i, k in enumerate(own):
        m = (L - k.shape[0]) // 2
        inner = dev[i, m:L - m, m:L - m].numpy()
        assert np.allclose(inner, k.astype(np.float32)) and abs(float(dev[i].sum()) - float(k.sum())) < 1e-6

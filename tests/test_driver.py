"""Host side of the evaluation loop (dynavsr_b200/driver.py): PSNR / SSIM restatements against the numbers the reference's
own test driver produced (tests/golden/driver_*.npz), the writer thread, the CSV."""
import os
import threading
import time

import numpy as np
import pytest
import torch

from util import gold


def _gt_image(g):
    from dynavsr_b200.driver import _host_image
    return _host_image(torch.from_numpy(g['gt']))


@pytest.mark.parametrize('tag', ['adam1_cb', 'sgd2_l2'])
def test_psnr_and_ssim_match_reference_driver_numbers(tag):
    from dynavsr_b200.driver import psnr_from_sse, ssim_u8
    g = gold('driver_%s.npz' % tag)
    img, gt = g['image'], _gt_image(g)
    sse = int(((img.astype(np.int64) - gt.astype(np.int64)) ** 2).sum())
    assert psnr_from_sse(sse, img.size) == pytest.approx(float(g['psnr_adapted']), abs=1e-9)
    assert ssim_u8(img, gt) == pytest.approx(float(g['ssim_adapted']), abs=1e-9)
    assert psnr_from_sse(0, 10) == float('inf')
    assert ssim_u8(img[..., 0], img[..., 0]) == pytest.approx(1.0)
    with pytest.raises(ValueError):
        ssim_u8(img, img[:-1])


def test_frame_writer_orders_jobs_and_propagates_errors():
    from dynavsr_b200.driver import FrameWriter, _PinnedRing
    seen = []
    w = FrameWriter(depth=4)
    for i in range(20):
        w.put(None, i, lambda v: (time.sleep(0.001), seen.append(v)))
    w.close()
    assert seen == list(range(20))
    w.close()                                   # idempotent

    w = FrameWriter(depth=4)
    ran = []

    def job(v):
        if v == 2:
            raise IOError('disk full')
        ran.append(v)

    with pytest.raises(IOError):
        for i in range(50):                     # the producer learns about the failure without dead-locking
            w.put(None, i, job)
        w.close()
    assert ran == [0, 1]                        # nothing runs after the first failure


def test_pinned_ring_blocks_until_release():
    from dynavsr_b200.driver import _PinnedRing
    ring = _PinnedRing.__new__(_PinnedRing)     # no pinned allocation on the CPU box: exercise the slot accounting only
    import queue
    ring.buffers, ring.free = [None, None], queue.Queue()
    for i in range(2):
        ring.free.put(i)
    a, b = ring.acquire(), ring.acquire()
    got = []
    t = threading.Thread(target=lambda: got.append(ring.acquire()))
    t.start()
    time.sleep(0.05)
    assert not got                              # both slots are out
    ring.release(a)
    t.join(2)
    assert got == [a] and b != a


def test_csv_and_summary(tmp_path):
    from collections import OrderedDict
    from dynavsr_b200.driver import COLUMNS, summary, write_csv
    rows = OrderedDict([('calendar/00000000', [28.0, 31.0, 0.91, 0.93]), ('calendar/00000001', [30.0, 33.0, 0.92, 0.94]),
                        ('city/00000000', [25.0, 27.0, float('nan'), 0.8])])
    path = tmp_path / 'psnr_update.csv'
    write_csv(rows, str(path))
    import pandas as pd
    df = pd.read_csv(str(path), index_col=0)
    assert list(df.columns) == COLUMNS and list(df.index) == list(rows)
    assert df.loc['calendar/00000001', 'PSNR_Ours'] == 33.0 and np.isnan(df.loc['city/00000000', 'SSIM_Bicubic'])
    s = summary(rows)
    assert s['calendar'][:2] == [29.0, 32.0] and s['city'][1] == 27.0
    assert s['__all__'][1] == pytest.approx((32.0 + 27.0) / 2)          # mean of per-folder means (test_dynavsr.py:308-363)


# ---------------------------------------------------------------------------------------------------------------
# evaluate(): host logic (sharding, slot accounting, ordering, table assembly) with stand-ins for the device pieces.
# The numerical side of the same function is covered on the GPU (tests/test_driver_gpu.py).
class _FakeOps(object):
    @staticmethod
    def to_nhwc(x):
        return x.permute(0, 2, 3, 1).contiguous()

    @staticmethod
    def frame_to_u8(frame, out=None, ref=None, sse=None, bgr=False):
        img = (frame.detach().clamp(0, 1) * 255.0).round().to(torch.uint8)
        if ref is not None and sse is not None:
            sse += ((img.long() - ref.long()) ** 2).sum()
        return img


class _FakeStream(object):
    def wait_stream(self, other):
        pass


class _FakeEvent(object):
    waited = 0

    def record(self):
        pass

    def synchronize(self):
        _FakeEvent.waited += 1


class _FakeEngine(object):
    """'Adaptation' = nearest-neighbour enlargement of the centre frame, plus a per-engine offset to tell engines apart."""

    class _Net(object):
        scale = 2

    def __init__(self, bias=0.0):
        self.netG, self.bias, self.calls = self._Net(), bias, 0

    def adapt_and_infer_nhwc(self, frames):
        self.calls += 1
        c = frames[frames.shape[0] // 2]
        return (c.repeat_interleave(2, 0).repeat_interleave(2, 1) + self.bias)[None]


class _FakePool(object):
    def __init__(self, n):
        self.engines, self.streams, self._next, self.scale, self.joined = [_FakeEngine() for _ in range(n)], [_FakeStream() for _ in range(n)], 0, 2, 0

    def submit(self, fn, pipeline=None):
        i = self._next
        self._next = (i + 1) % len(self.engines)
        return i, fn(self.engines[i])

    def join(self):
        self.joined += 1


@pytest.fixture()
def fake_device(monkeypatch):
    import contextlib
    from dynavsr_b200 import driver
    monkeypatch.setattr(driver, 'ops', _FakeOps)
    monkeypatch.setattr(driver, '_DEVICE', 'cpu')
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a: _FakeStream())
    monkeypatch.setattr(torch.cuda, 'stream', lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, 'Event', _FakeEvent)
    monkeypatch.setattr(torch.Tensor, 'record_stream', lambda self, s: None)
    return driver


def _items(n, with_gt=True, h=8, w=12):
    g = torch.Generator().manual_seed(3)
    out = []
    for i in range(n):
        lq = torch.rand(1, 5, 3, h, w, generator=g)
        d = {'LQs': lq, 'folder': ['a' if i < 3 else 'b'], 'idx': ['%d/%d' % (i, n)]}
        if with_gt:
            gt = lq[0, 2].repeat_interleave(2, 1).repeat_interleave(2, 2)                   # exactly what _FakeEngine makes
            d['GT'] = gt[None, None].expand(1, 5, -1, -1, -1) + (0.1 if i == 1 else 0.0)     # frame 1: a known error
        out.append(d)
    return out


@pytest.mark.parametrize('pooled', [False, True])
def test_evaluate_host_logic(fake_device, tmp_path, pooled):
    driver = fake_device
    written = []
    engine = _FakePool(3) if pooled else _FakeEngine()
    rows = driver.evaluate(engine, _items(5), save_dir=str(tmp_path), ring_slots=2,
                           sink=lambda path, img: written.append((path, img.copy())))
    assert list(rows) == ['a/00000000', 'a/00000001', 'a/00000002', 'b/00000003', 'b/00000004']
    assert [os.path.relpath(p, str(tmp_path)) for p, _ in written] == [
        'a/DynaVSR/00000000.png', 'a/DynaVSR/00000001.png', 'a/DynaVSR/00000002.png', 'b/DynaVSR/00000003.png',
        'b/DynaVSR/00000004.png']                                   # in order, although only two pinned slots exist
    assert all(img.shape == (16, 24, 3) and img.dtype == np.uint8 for _, img in written)
    for name, (p0, p1, s0, s1) in rows.items():
        assert np.isnan(p0) and np.isnan(s0)                        # no baseline model given
        if name == 'a/00000001':
            assert 19.0 < p1 < 21.5 and s1 < 1.0                    # GT is 0.1 off (clamped at 1): ~20 dB
        else:
            assert p1 == float('inf') and s1 == pytest.approx(1.0)
    if pooled:
        assert [e.calls for e in engine.engines] == [2, 2, 1] and engine.joined == 1     # round robin over the pipelines
    # sharding: rank r of 2 takes items r, r + 2, ...; the two tables partition the full one
    r0 = driver.evaluate(engine, _items(5), rank=0, world_size=2, compute_ssim=False)
    r1 = driver.evaluate(engine, _items(5), rank=1, world_size=2, compute_ssim=False)
    assert list(r0) == ['a/00000000', 'a/00000002', 'b/00000004'] and list(r1) == ['a/00000001', 'b/00000003']
    assert r1['a/00000001'][1] == rows['a/00000001'][1]
    # no ground truth ('demo' mode): images are written, every metric is NaN
    demo = driver.evaluate(engine, _items(2, with_gt=False), with_GT=False, save_dir=str(tmp_path / 'demo'), sink=lambda p, i: None)
    assert all(np.isnan(v).all() for v in demo.values()) and len(demo) == 2


def test_evaluate_baseline_column_and_sink_errors(fake_device, tmp_path):
    driver = fake_device

    class Baseline(object):
        def forward_nhwc(self, frames, B, N):
            return _FakeEngine(bias=0.05).adapt_and_infer_nhwc(frames)

    rows = driver.evaluate(_FakeEngine(), _items(3), baseline_netG=Baseline())
    for name, (p0, p1, s0, s1) in rows.items():
        if name != 'a/00000001':                                     # (frame 1 carries the deliberate GT error)
            assert 24.0 < p0 < 28.5 and s0 < 1.0 and p1 > p0         # the 'baseline' is 0.05 off everywhere: ~26 dB
    with pytest.raises(IOError):                                     # a failing sink surfaces in the caller, not silently
        driver.evaluate(_FakeEngine(), _items(4), save_dir=str(tmp_path), ring_slots=1,
                        sink=lambda p, i: (_ for _ in ()).throw(IOError('disk full')))
    # clips of different sizes get their own pinned buffers
    mixed = _items(2) + _items(1, h=12, w=8)
    mixed[2]['idx'], mixed[2]['folder'] = ['7/9'], ['c']
    rows = driver.evaluate(_FakeEngine(), mixed, compute_ssim=False)
    assert list(rows) == ['a/00000000', 'a/00000001', 'c/00000007']


def test_frame_writer_releases_slots_of_skipped_items_after_an_error():
    """after the first failing job the writer skips the remaining items; their pinned-ring slots must still come
    back, or the producer blocks in ring.acquire() instead of seeing the error at close()."""
    import torch
    from dynavsr_b200 import driver
    ring = driver._PinnedRing((4, 4, 3), 2)
    writer = driver.FrameWriter(depth=8)
    ran = []

    def submit(i, fail):
        slot = ring.acquire()                       # would dead-lock on the third item if slots leaked
        done = []

        def release():
            if not done:
                done.append(1)
                ring.release(slot)

        def job(img):
            try:
                if fail:
                    raise ValueError('boom %d' % i)
                ran.append(i)
            finally:
                release()
        job.release = release
        writer._q.put((None, torch.zeros(1), job))

    for i in range(6):
        submit(i, fail=(i == 1))
    with pytest.raises(ValueError, match='boom 1'):
        writer.close()
    assert ran == [0]                               # items after the failure were skipped ...
    assert ring.free.qsize() == 2                   # ... and every slot is back in the ring

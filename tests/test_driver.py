"""Host side of the evaluation loop (dynavsr_b200/driver.py): PSNR / SSIM restatements against the numbers the reference's
own test driver produced (tests/golden/driver_*.npz), the writer thread, the CSV."""
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

"""Test-time evaluation loop: the host side of codes/test_dynavsr.py:155-365 around the adaptation engines.

Per clip the reference driver (1) runs the baseline model, (2) deep-copies both networks, builds an optimiser, adapts for K
steps, (3) runs the adapted model, (4) converts both results to 8-bit images on the host (``tensor2img``), computes PSNR and
SSIM against the ground truth, writes a PNG and rewrites a CSV -- all on one thread, every item a device synchronisation.
Here the engines already own (2) (flat-buffer restore, persistent optimiser state, CUDA graphs); this module keeps (1), (3)
and (4) off the critical path:

  * the 8-bit image is made on the GPU by the kernel that also accumulates the exact integer squared error against the
    ground-truth image (``ops.frame_to_u8``), so PSNR needs no host copy of a float frame and no per-frame sync -- the
    squared errors of the whole run come back in ONE device-to-host copy at the end;
  * images travel device -> pinned host memory asynchronously (2.7 MB per 720p frame instead of 10.8 MB of float32) and a
    writer thread waits on the copy's event, encodes the PNG and computes SSIM while the GPU works on the next frames;
  * frames are independent (test_dynavsr.py:208), so with an ``AdaptationPool`` several are in flight at once, and across
    processes they are sharded ``range(rank, n, world)`` exactly like the reference's validation (train_dynavsr.py:509).

``evaluate`` returns the table the reference writes to ``psnr_update.csv`` (same index and column names).
"""
import math
import os
import queue
import threading
from collections import OrderedDict

import numpy as np
import torch

from . import ops

COLUMNS = ['PSNR_Bicubic', 'PSNR_Ours', 'SSIM_Bicubic', 'SSIM_Ours']        # test_dynavsr.py:118
_DEVICE = 'cuda'            # the loop's tensors live here (host-logic tests substitute stand-ins for the device pieces)


def psnr_from_sse(sse, n):
    """calculate_psnr (utils/util.py:262-269) from the integer squared error of two 8-bit images of ``n`` bytes."""
    return float('inf') if sse == 0 else 20.0 * math.log10(255.0 / math.sqrt(sse / float(n)))


def _gaussian_window(size=11, sigma=1.5):
    x = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return k / k.sum()


def _blur_valid(img, k):
    """Separable Gaussian filtering of [H, W, C], 'valid' part only (the reference crops [5:-5, 5:-5] after filter2D)."""
    n = len(k)
    H, W = img.shape[0], img.shape[1]
    rows = sum(k[i] * img[i:H - n + 1 + i] for i in range(n))
    return sum(k[i] * rows[:, i:W - n + 1 + i] for i in range(n))


def ssim_u8(img1, img2):
    """utils/util.py:272-313 (calculate_ssim): 11 x 11 Gaussian window (sigma 1.5), C1 = (0.01*255)^2, C2 = (0.03*255)^2,
    valid region, mean over pixels, then over the colour channels.  Inputs: HWC (or HW) arrays in [0, 255]."""
    if img1.shape != img2.shape:
        raise ValueError('Input images must have the same dimensions.')
    a = np.asarray(img1, dtype=np.float64)
    b = np.asarray(img2, dtype=np.float64)
    if a.ndim == 2:
        a, b = a[..., None], b[..., None]
    if a.ndim != 3:
        raise ValueError('Wrong input image dimensions.')
    k = _gaussian_window()
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    mu1, mu2 = _blur_valid(a, k), _blur_valid(b, k)
    s1 = _blur_valid(a * a, k) - mu1 * mu1
    s2 = _blur_valid(b * b, k) - mu2 * mu2
    s12 = _blur_valid(a * b, k) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2))
    return float(m.mean(axis=(0, 1)).mean())


def write_png(path, rgb):
    """Default image sink: PNG through OpenCV (the reference uses imageio.imwrite on the same RGB array)."""
    import cv2
    if not cv2.imwrite(path, np.ascontiguousarray(rgb[..., ::-1])):
        raise IOError('could not write %s' % path)


class FrameWriter(object):
    """Consumes (event, pinned image, job) items on a thread: waits for the device-to-host copy, then runs ``job(image)``
    (PNG encoding, SSIM, ...).  ``close()`` drains the queue and re-raises the first error a job raised."""

    def __init__(self, depth=16):
        self._q = queue.Queue(maxsize=depth)
        self._err = None
        self._t = threading.Thread(target=self._run, name='dvsr-frame-writer', daemon=True)
        self._t.start()

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            event, image, job = item
            try:
                if self._err is None:
                    if event is not None:
                        event.synchronize()
                    job(image)
                    continue
            except Exception as e:           # keep draining so producers never block on a dead consumer
                if self._err is None:
                    self._err = e
            # the job did not run to completion (an earlier error, or this one): give its resources back -- a pinned-ring slot
            # that is never released would leave the producer blocked in ring.acquire() instead of seeing the error
            release = getattr(job, 'release', None)
            if release is not None:
                try:
                    release()
                except Exception:
                    pass

    def put(self, event, image, job):
        if self._err is not None:
            self.close()
        self._q.put((event, image, job))

    def close(self):
        if self._t.is_alive():
            self._q.put(None)
            self._t.join()
        if self._err is not None:
            err, self._err = self._err, None
            raise err


class _PinnedRing(object):
    """Pinned host image buffers reused round-robin; a slot is handed out again only after its job has finished."""

    def __init__(self, shape, slots):
        pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
        self.buffers = [pin(torch.empty(shape, dtype=torch.uint8)) for _ in range(slots)]
        self.free = queue.Queue()
        for i in range(slots):
            self.free.put(i)

    def acquire(self):
        return self.free.get()

    def release(self, i):
        self.free.put(i)


def _frame_name(folder, idx):
    return '{}/{:08d}'.format(folder, idx)                  # test_dynavsr.py:289


def _first(v):
    return v[0] if isinstance(v, (list, tuple)) else v


class _SseTable(object):
    """Device-side int64 accumulators, two per frame (baseline, adapted).  Grows by whole blocks so that kernels already
    enqueued on other streams keep writing into storage that stays alive and in place."""

    BLOCK = 2048

    def __init__(self):
        self.blocks = []

    def cell(self, i):
        b, o = divmod(i, self.BLOCK)
        while b >= len(self.blocks):
            self.blocks.append(torch.zeros(self.BLOCK, dtype=torch.int64, device=_DEVICE))
        return self.blocks[b][o:o + 1]

    def tolist(self):
        return torch.cat(self.blocks).cpu().tolist() if self.blocks else []


def _host_image(t):
    """tensor2img (utils/util.py:112-142, mode='rgb') of a host float tensor [3, H, W] -> uint8 [H, W, 3]."""
    return (t.detach().float().cpu().clamp(0, 1) * 255.0).round().to(torch.uint8).permute(1, 2, 0).contiguous().numpy()


def evaluate(engine, loader, baseline_netG=None, with_GT=True, save_dir=None, compute_ssim=True, rank=0, world_size=1,
             sink=write_png, ring_slots=8):
    """Run the reference's test loop over ``loader`` and return an OrderedDict ``name -> [PSNR_Bicubic, PSNR_Ours,
    SSIM_Bicubic, SSIM_Ours]`` (NaN where not computed) for this rank's frames.

    engine         an ``InnerLoopAdapter`` or an ``AdaptationPool`` (several frames in flight)
    loader         iterable of the reference datasets' items (video_test_dataset_int.py:212-239): ``LQs`` [1, N, 3, h, w],
                   ``GT`` [1, N, 3, sh, sw] (when ``with_GT``), ``folder``, ``idx`` ('i/n')
    baseline_netG  the un-adapted baseline EDVR (``path.bicubic_G``, test_dynavsr.py:197-205) or None to skip it
    save_dir       where ``<folder>/DynaVSR/<idx:08d>.png`` is written (:158-165, :283); None = images are not written
    sink           ``sink(path, rgb_uint8_hwc)``; runs on the writer thread
    """
    pool = engine if hasattr(engine, 'submit') else None
    scale = engine.scale if pool is not None else engine.netG.scale
    writer = FrameWriter(depth=2 * ring_slots)
    table = _SseTable()
    records = []                       # (name, bytes per image)
    ssim_vals = {}
    lock = threading.Lock()
    rings = {}                         # pinned image buffers per image shape (clips of one dataset share a shape)
    do_base = baseline_netG is not None and with_GT

    def emit(ring, image_dev, event_stream, job):
        """Asynchronous device -> pinned host copy of an 8-bit image on ``event_stream``; ``job(ndarray)`` runs on the writer
        thread once the copy has landed, then the pinned slot is recycled."""
        slot = ring.acquire()                     # blocks only while every pinned buffer is still being consumed
        host = ring.buffers[slot]
        with torch.cuda.stream(event_stream):
            host.copy_(image_dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()

        released = []

        def release():
            if not released:
                released.append(True)
                ring.release(slot)

        def wrapped(img):
            try:
                job(img.numpy())
            finally:
                release()
        wrapped.release = release                 # FrameWriter calls it for items it has to skip after an error
        writer.put(ev, host, wrapped)

    try:
        for item_idx, data in enumerate(loader):
            if item_idx % world_size != rank:
                continue
            folder = _first(data['folder'])
            idx_d = int(str(_first(data['idx'])).split('/')[0])
            name = _frame_name(folder, idx_d)
            lq = data['LQs']
            assert lq.size(0) == 1                                                      # :186
            _, N, C, h, w = lq.shape
            shape = (h * scale, w * scale, C)
            if shape not in rings:
                rings[shape] = _PinnedRing(shape, ring_slots)
            ring = rings[shape]
            i = len(records)
            records.append((name, h * scale * w * scale * C))
            frames = ops.to_nhwc(lq.to(_DEVICE, non_blocking=True).reshape(N, C, h, w))
            gt_u8, gt_img = None, [None]
            if with_GT:
                gt = data['GT'][0, N // 2]                                              # :183
                gt_u8 = ops.frame_to_u8(ops.to_nhwc(gt.unsqueeze(0).to(_DEVICE, non_blocking=True))[0])
            gt_ready = None
            if with_GT and compute_ssim and gt.is_cuda:
                gt_ready = torch.cuda.Event()          # a device-resident clip: the writer thread reads ``gt`` on ITS stream,
                gt_ready.record()                      # so it first waits until this stream has produced the frame

            def ssim_job(kind, gt=gt if with_GT else None, gt_img=gt_img, name=name, folder=folder, idx_d=idx_d,
                         gt_ready=gt_ready):
                def job(img):
                    if kind == 1 and save_dir is not None:                              # :158-165, :283
                        d = os.path.join(save_dir, folder, 'DynaVSR')
                        os.makedirs(d, exist_ok=True)
                        sink(os.path.join(d, '{:08d}.png'.format(idx_d)), img)
                    if compute_ssim and gt is not None:
                        if gt_img[0] is None:       # jobs run one at a time on the writer thread
                            if gt_ready is not None:
                                gt_ready.synchronize()
                            gt_img[0] = _host_image(gt)
                        v = ssim_u8(img, gt_img[0])
                        with lock:
                            ssim_vals[(name, kind)] = v
                return job

            cur = torch.cuda.current_stream()
            if do_base:                                                                 # :197-205
                with torch.no_grad():
                    base = baseline_netG.forward_nhwc(frames, 1, N)[0]
                base_u8 = ops.frame_to_u8(base, ref=gt_u8, sse=table.cell(2 * i))
                if compute_ssim:
                    emit(ring, base_u8, cur, ssim_job(0))
            cell = table.cell(2 * i + 1) if with_GT else None

            def run(eng, frames=frames, gt_u8=gt_u8, cell=cell):
                hr = eng.adapt_and_infer_nhwc(frames)[0]                                # :208-277
                return ops.frame_to_u8(hr, ref=gt_u8, sse=cell)

            if pool is not None:
                stream = pool.streams[pool._next]
                frames.record_stream(stream)
                if gt_u8 is not None:
                    gt_u8.record_stream(stream)
                _, img = pool.submit(run)
                img.record_stream(cur)
                emit(ring, img, stream, ssim_job(1))
            else:
                emit(ring, run(engine), cur, ssim_job(1))
        if pool is not None:
            pool.join()
        writer.close()
        sse_host = table.tolist()                  # the run's only blocking device-to-host read
    finally:
        try:
            writer.close()
        except Exception:
            pass
    nan = float('nan')
    rows = OrderedDict()
    for i, (name, n) in enumerate(records):
        rows[name] = [psnr_from_sse(sse_host[2 * i], n) if do_base else nan,
                      psnr_from_sse(sse_host[2 * i + 1], n) if with_GT else nan,
                      ssim_vals.get((name, 0), nan), ssim_vals.get((name, 1), nan)]
    return rows


def gather_rows(rows, dst=0):
    """All ranks' tables merged on rank ``dst`` (the reference reduces zero-filled per-folder vectors,
    train_dynavsr.py:722-728; names are unique per frame, so a gather of (name, values) pairs is equivalent)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return rows
    parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(list(rows.items()), parts, dst=dst)
    if dist.get_rank() != dst:
        return None
    merged = OrderedDict()
    for part in parts:
        merged.update(part)
    return OrderedDict(sorted(merged.items()))


def write_csv(rows, path):
    """``psnr_update.csv`` as the reference's pandas frame prints it (test_dynavsr.py:118, :289-301)."""
    with open(path, 'w') as f:
        f.write(',' + ','.join(COLUMNS) + '\n')
        for name, vals in rows.items():
            f.write(name + ',' + ','.join('' if v != v else repr(float(v)) for v in vals) + '\n')


def summary(rows):
    """Per-folder and overall averages, the numbers the reference prints at the end (:308-363)."""
    per_folder = OrderedDict()
    for name, vals in rows.items():
        per_folder.setdefault(name.rsplit('/', 1)[0], []).append(vals)
    out = OrderedDict()
    for folder, vals in per_folder.items():
        out[folder] = [float(np.mean([v[c] for v in vals])) for c in range(len(COLUMNS))]
    out['__all__'] = [float(np.mean([v[c] for v in out.values()])) for c in range(len(COLUMNS))] if out else []
    return out

"""Clip store and frame-window assembly: the ``cache_data`` path of codes/data/meta_learner/video_test_dataset_int.py:150-246
(every clip decoded once, items assembled by ``index_select`` of N frame indices around the centre) with the clips resident
on the GPU, so a window is a device-side gather of N small frames instead of a host tensor that is re-uploaded N times per
clip.  Items have the keys and shapes a ``DataLoader(batch_size=1)`` over the reference dataset yields -- what
``driver.evaluate`` and ``test_dynavsr.py:155-190`` consume.

Host logic only (index lists, cropping rule, item layout); arithmetic on the frames (degradation) is
``degradation.Degradation``.
"""
from collections import OrderedDict

import torch

PADDINGS = ('replicate', 'reflection', 'new_info', 'circle')


def index_generation(crt_i, max_n, N, padding='reflection'):
    """Frame indices of the N-frame window centred on ``crt_i`` in a clip of ``max_n`` frames (codes/data/util.py:114-160).
    Positions that fall off either end are filled according to ``padding``; for crt_i = 0, N = 5:
    replicate [0, 0, 0, 1, 2], reflection [2, 1, 0, 1, 2], new_info [4, 3, 0, 1, 2], circle [3, 4, 0, 1, 2]."""
    if padding not in PADDINGS:
        raise ValueError('Wrong padding mode')
    last, half = max_n - 1, N // 2
    lo, hi = crt_i - half, crt_i + half
    before = {'replicate': lambda i: 0, 'reflection': lambda i: -i, 'new_info': lambda i: hi - i, 'circle': lambda i: N + i}[padding]
    after = {'replicate': lambda i: last, 'reflection': lambda i: 2 * last - i, 'new_info': lambda i: lo - (i - last),
             'circle': lambda i: i - N}[padding]
    return [before(i) if i < 0 else (after(i) if i > last else i) for i in range(lo, hi + 1)]


class ResidentClips(object):
    """Clips kept on ``device`` as [T, 3, H, W] float32 in [0, 1]; ``store[k]`` / iteration yield one item per output frame.

    n_frames  N, the window length (``datasets.val.N_frames``)
    padding   index rule at the clip ends (``datasets.val.padding``)
    scale     SR factor: ground truth is ``scale`` x the LR size; the adaptation runs EDVR on LR / scale, so LR height and
              width are cropped to multiples of 4 * scale exactly as video_test_dataset_int.py:185-189 crops SLQ / LQ / GT
    """

    def __init__(self, n_frames=5, padding='new_info', scale=4, device=None):
        if padding not in PADDINGS:
            raise ValueError('Wrong padding mode')
        self.N, self.padding, self.scale = n_frames, padding, scale
        self.device = torch.device(device) if device is not None else torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        self.clips = OrderedDict()            # folder -> dict(LQ=..., GT=... or None, SLQ=... or None)
        self._index = []                      # (folder, frame index)

    # ------------------------------------------------------------------ filling
    def add(self, folder, lq, gt=None, slq=None):
        """lq [T, 3, h, w]; gt [T, 3, s*h, s*w] or None (the reference's 'demo' mode); slq [T, 3, h/s, w/s] or None."""
        if folder in self.clips:
            raise ValueError('clip [%s] was already added' % folder)
        T, _, h, w = lq.shape
        s = self.scale
        if gt is not None and (gt.shape[0] != T or gt.shape[-2] != s * h or gt.shape[-1] != s * w):
            raise ValueError('GT %s does not match LQ %s at scale %d' % (tuple(gt.shape), tuple(lq.shape), s))
        if slq is not None and (slq.shape[0] != T or slq.shape[-2] * s != h or slq.shape[-1] * s != w):
            raise ValueError('SuperLQ %s does not match LQ %s at scale %d' % (tuple(slq.shape), tuple(lq.shape), s))
        hs, ws = h // s, w // s                                   # super-LR size; cropped to multiples of 4
        hs, ws = hs - hs % 4, ws - ws % 4
        if hs == 0 or ws == 0:
            raise ValueError('clip [%s] is too small: %dx%d at scale %d' % (folder, h, w, s))
        put = lambda t, f: None if t is None else t[..., :hs * f, :ws * f].to(self.device, torch.float32).contiguous()
        self.clips[folder] = dict(LQ=put(lq, s), GT=put(gt, s * s), SLQ=put(slq, 1))
        self._index.extend((folder, i) for i in range(T))
        return self

    def add_degraded(self, folder, gt, degradation):
        """LQ (and SuperLQ) made from ground-truth frames with ``degradation.Degradation.apply`` (the synthetic protocol:
        vsrbase.py:184-192 applies the same kernel once for LQ and twice for SuperLQ)."""
        s = self.scale
        f = 4 * s * s
        gt = gt[..., :gt.shape[-2] - gt.shape[-2] % f, :gt.shape[-1] - gt.shape[-1] % f].to(self.device, torch.float32)
        lq = degradation.apply(gt)
        return self.add(folder, lq, gt=gt, slq=degradation.apply(lq))

    # ------------------------------------------------------------------ items
    def __len__(self):
        return len(self._index)

    def window_indices(self, index):
        folder, i = self._index[index]
        return index_generation(i, self.clips[folder]['LQ'].shape[0], self.N, padding=self.padding)

    def item(self, index):
        """The reference dataset's ``__getitem__`` (video_test_dataset_int.py:212-239), no batch dimension."""
        folder, i = self._index[index]
        clip = self.clips[folder]
        T = clip['LQ'].shape[0]
        sel = torch.tensor(index_generation(i, T, self.N, padding=self.padding), device=self.device)
        half = self.N // 2
        out = {'LQs': clip['LQ'].index_select(0, sel), 'folder': folder, 'idx': '{}/{}'.format(i, T),
               'border': 1 if (i < half or i >= T - half) else 0}
        if clip['GT'] is not None:
            out['GT'] = clip['GT'].index_select(0, sel)
        if clip['SLQ'] is not None:
            out['SuperLQs'] = clip['SLQ'].index_select(0, sel)
        return out

    def __getitem__(self, index):
        """As ``DataLoader(batch_size=1)`` collates an item: tensors gain a leading 1, strings become 1-element lists."""
        it = self.item(index)
        return {k: (v.unsqueeze(0) if torch.is_tensor(v) else ([v] if isinstance(v, str) else torch.tensor([v])))
                for k, v in it.items()}

    def __iter__(self):
        return (self[i] for i in range(len(self)))

"""Clip store / window assembly (dynavsr_b200/clips.py) against the reference's index table
(tests/golden/index_generation.json, tabulated from the unmodified codes/data/util.py by oracle/make_golden_index.py)."""
import json
import os

import pytest
import torch

from util import GOLD


def test_index_generation_matches_reference_table():
    from dynavsr_b200.clips import index_generation
    table = json.load(open(os.path.join(GOLD, 'index_generation.json')))
    assert len(table) == 580
    for key, want in table.items():
        padding, N, max_n, crt = key.split('/')
        assert index_generation(int(crt), int(max_n), int(N), padding=padding) == want, key
    assert index_generation(0, 100, 5, 'new_info') == [4, 3, 0, 1, 2]           # the docstring examples of util.py:121-125
    assert index_generation(0, 100, 5, 'circle') == [3, 4, 0, 1, 2]
    with pytest.raises(ValueError):
        index_generation(0, 10, 5, 'mirror')


def _clip(T, h, w, s, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(T, 3, h, w, generator=g), torch.rand(T, 3, h * s, w * s, generator=g),
            torch.rand(T, 3, h // s, w // s, generator=g))


def test_resident_clips_items_follow_the_reference_dataset_layout():
    from dynavsr_b200.clips import ResidentClips, index_generation
    store = ResidentClips(n_frames=5, padding='new_info', scale=2, device='cpu')
    lq_a, gt_a, slq_a = _clip(7, 16, 24, 2, 1)
    lq_b, gt_b, slq_b = _clip(6, 18, 28, 2, 2)          # 18x28 -> super-LR 9x14 -> cropped to 8x12 (multiples of 4), LQ 16x24
    store.add('calendar', lq_a, gt_a, slq_a).add('city', lq_b, gt_b, slq_b)
    assert len(store) == 13
    it = store.item(1)                                    # frame 1 of 'calendar'
    sel = index_generation(1, 7, 5, 'new_info')
    assert it['folder'] == 'calendar' and it['idx'] == '1/7' and it['border'] == 1
    assert torch.equal(it['LQs'], lq_a[sel]) and torch.equal(it['GT'], gt_a[sel]) and torch.equal(it['SuperLQs'], slq_a[sel])
    assert store.item(3)['border'] == 0 and store.item(6)['border'] == 1
    it = store.item(7 + 5)                                # last frame of 'city', cropped
    sel = index_generation(5, 6, 5, 'new_info')
    assert it['idx'] == '5/6' and it['LQs'].shape == (5, 3, 16, 24) and it['GT'].shape == (5, 3, 32, 48)
    assert it['SuperLQs'].shape == (5, 3, 8, 12)
    assert torch.equal(it['LQs'], lq_b[sel][..., :16, :24]) and torch.equal(it['GT'], gt_b[sel][..., :32, :48])
    # batched form = what DataLoader(batch_size=1) hands to test_dynavsr.py:155-190
    b = store[1]
    assert b['LQs'].shape == (1, 5, 3, 16, 24) and b['folder'] == ['calendar'] and b['idx'] == ['1/7']
    assert int(b['idx'][0].split('/')[0]) == 1 and int(b['border']) == 1
    assert [d['idx'][0] for d in store][:3] == ['0/7', '1/7', '2/7']


def test_resident_clips_argument_checks():
    from dynavsr_b200.clips import ResidentClips
    store = ResidentClips(scale=4, device='cpu')
    lq, gt, slq = _clip(5, 32, 32, 4, 3)
    store.add('a', lq)                                    # 'demo' mode: no ground truth
    assert 'GT' not in store.item(0) and 'SuperLQs' not in store.item(0)
    with pytest.raises(ValueError):
        store.add('a', lq)                                # same folder twice
    with pytest.raises(ValueError):
        store.add('b', lq, gt[:, :, :-4])                 # GT not scale x LQ
    with pytest.raises(ValueError):
        store.add('c', lq, gt, slq[:4])                   # frame count mismatch
    with pytest.raises(ValueError):
        store.add('d', lq[..., :8, :8])                   # smaller than one 4 x scale block
    with pytest.raises(ValueError):
        ResidentClips(padding='mirror')

"""oracle/make_golden_index.py -- TEST INFRASTRUCTURE ONLY; run in the build container.

Tabulates the UNMODIFIED reference ``index_generation`` (codes/data/util.py:114-160) over every padding mode, window
length 1-7 and clip length up to 9 -> tests/golden/index_generation.json (the reference ships no test for it).

    python -m oracle.make_golden_index
"""
import json
import os
import sys
import types

REF = os.environ.get('DVSR_REFERENCE', '/root/reference/codes')
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main():
    sys.path.insert(0, REF)
    for name in ('imageio', 'lmdb'):
        sys.modules.setdefault(name, types.ModuleType(name))
    import data.util as U
    table = {}
    for padding in ('replicate', 'reflection', 'new_info', 'circle'):
        for N in (1, 3, 5, 7):
            for max_n in range(max(N, 2), 10):
                for crt in range(max_n):
                    table['%s/%d/%d/%d' % (padding, N, max_n, crt)] = U.index_generation(crt, max_n, N, padding=padding)
    with open(os.path.join(GOLD, 'index_generation.json'), 'w') as f:
        json.dump(table, f, separators=(',', ':'))
    print(len(table), 'cases')


if __name__ == '__main__':
    main()

"""Model factory with the reference's names (codes/models/__init__.py:5-37): ``create_model(opt)`` returns one wrapper, or
for ``'a+b'`` a list in that order -- ``'video_base+lrimgestimator'`` -> ``[VideoBaseModel, LRimgestimator_Model]``, which is
what test_dynavsr.py:105-109 and train_dynavsr.py:161-165 unpack.  Only the two wrappers on the DynaVSR hot path exist here."""
import importlib
import logging

logger = logging.getLogger('base')

_WRAPPERS = {'video_base': ('Video_base_model', 'VideoBaseModel'),
             'lrimgestimator': ('LRestimator_model', 'LRimgestimator_Model')}
_OUTSIDE_HOT_PATH = ('sr', 'srgan', 'classifier', 'estimator')


def _build(name, opt):
    if name not in _WRAPPERS:
        why = 'is outside the DynaVSR hot path (SURVEY.md section 2 rows 20-21).' if name in _OUTSIDE_HOT_PATH \
            else 'not recognized.'
        raise NotImplementedError('Model [{:s}] {}'.format(name, why))
    module, cls = _WRAPPERS[name]
    wrapper = getattr(importlib.import_module('.' + module, __name__), cls)(opt)
    logger.info('Model [{:s}] is created.'.format(cls))
    return wrapper


def create_model(opt):
    spec = opt['model']
    built = [_build(name, opt) for name in spec.split('+')]
    return built if '+' in spec else built[0]

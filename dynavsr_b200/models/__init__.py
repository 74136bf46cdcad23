"""Mirror of codes/models/__init__.py:5-37: ``create_model(opt)`` with the reference's model names and the
``'video_base+lrimgestimator'`` -> ``[VideoBaseModel, LRimgestimator_Model]`` list order that test_dynavsr.py:105-109 and
train_dynavsr.py:161-165 unpack.  Only the two wrappers on the DynaVSR hot path exist here."""
import logging

logger = logging.getLogger('base')


def create_model(opt):
    models = opt['model']

    def _create(model):
        if model == 'video_base':
            from .Video_base_model import VideoBaseModel as M
        elif model == 'lrimgestimator':
            from .LRestimator_model import LRimgestimator_Model as M
        elif model in ('sr', 'srgan', 'classifier', 'estimator'):
            raise NotImplementedError('Model [{:s}] is outside the DynaVSR hot path (SURVEY.md section 2 rows 20-21).'.format(model))
        else:
            raise NotImplementedError('Model [{:s}] not recognized.'.format(model))
        m = M(opt)
        logger.info('Model [{:s}] is created.'.format(m.__class__.__name__))
        return m

    if '+' in models:
        return [_create(name) for name in models.split('+')]
    return _create(models)

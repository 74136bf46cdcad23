"""Mirror of codes/models/networks.py:8-45,60-72: option dict -> network, for the two networks of the hot path."""
from .archs import EDVR_arch, LRimg_estimator


def define_G(opt):
    opt_net = opt['network_G']
    which_model = opt_net['which_model_G']
    if which_model == 'EDVR':
        return EDVR_arch.EDVR(nf=opt_net['nf'], nframes=opt_net['nframes'], groups=opt_net['groups'],
                              front_RBs=opt_net['front_RBs'], back_RBs=opt_net['back_RBs'], center=opt_net['center'],
                              predeblur=bool(opt_net['predeblur']), HR_in=bool(opt_net['HR_in']),
                              w_TSA=opt_net['w_TSA'] if opt_net['w_TSA'] is not None else True, scale=opt['scale'])
    if which_model in ('MSRResNet', 'RRDBNet', 'DUF', 'TOF'):
        raise NotImplementedError('Generator model [{:s}] is outside the DynaVSR hot path (EDVR only).'.format(which_model))
    raise NotImplementedError('Generator model [{:s}] not recognized'.format(which_model))


def define_E(opt):
    opt_net = opt['network_E']
    which_model = opt_net['which_model_E']
    if which_model == 'MFDN':
        return LRimg_estimator.DirectKernelEstimatorVideo(in_nc=opt_net['in_nc'], nf=opt_net['nf'], scale=opt['scale'])
    if which_model == 'SFDN':
        return LRimg_estimator.DirectKernelEstimator_CMS(nf=opt_net['nf'])
    raise NotImplementedError('Estimator model [{:s}] not recognized'.format(which_model))

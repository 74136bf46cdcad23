"""Learning-rate schedules with restarts used by ``VideoBaseModel`` (``lr_scheme: MultiStepLR`` /
``CosineAnnealingLR_Restart``; reference: codes/models/lr_scheduler.py:8-64).  Closed forms in the iteration count t instead
of torch's chained ``_LRScheduler`` recursions (t counted in ``step()`` calls: the reference steps once per iteration,
base_model.py:51-53):

    multi-step:  lr(t) = initial_lr * w(last restart <= t) * gamma ** #{milestones m : last restart < m <= t}
    cosine:      lr(t) = eta_min + (A - eta_min) * (1 + cos(pi * (t - s) / T)) / 2
                 s = last restart <= t (0 before the first), T = the period that starts there,
                 A = initial_lr * w(s) down to the first trough (t - s <= T), initial_lr after it -- the reference's
                 trough-crossing branch (:56-60) re-enters the cosine with the unweighted base rate

Both are checked against traces of the unmodified reference classes (tests/golden/lr_schedules.json)."""
import math


class MultiStepLR_Restart(object):
    def __init__(self, optimizer, milestones, restarts=None, weights=None, gamma=0.1, clear_state=False, last_epoch=-1):
        self.optimizer = optimizer
        self.milestones = sorted(milestones or [])
        self.gamma = gamma
        self.clear_state = clear_state
        self.restarts = [v + 1 for v in (restarts if restarts else [0])]
        self.restart_weights = list(weights) if weights else [1]
        assert len(self.restarts) == len(self.restart_weights), 'restarts and their weights do not match.'
        for g in optimizer.param_groups:
            g.setdefault('initial_lr', g['lr'])
        self.last_epoch = last_epoch
        self.step()

    def _factor(self, t):
        start, w = 0, 1.0
        for r, rw in zip(self.restarts, self.restart_weights):
            if r <= t and r >= start:
                start, w = r, rw
        passed = sum(1 for m in self.milestones if start < m <= t)
        return w * self.gamma ** passed

    def get_lr(self):
        f = self._factor(self.last_epoch)
        return [g['initial_lr'] * f for g in self.optimizer.param_groups]

    def step(self, epoch=None):
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        if self.clear_state and self.last_epoch in self.restarts and hasattr(self.optimizer, 'flat'):
            fl = self.optimizer.flat
            if fl.m is not None:
                fl.m.zero_()
                fl.v.zero_()
        for g, lr in zip(self.optimizer.param_groups, self.get_lr()):
            g['lr'] = lr

    def state_dict(self):
        return {'last_epoch': self.last_epoch}

    def load_state_dict(self, sd):
        self.last_epoch = sd['last_epoch']


class CosineAnnealingLR_Restart(object):
    def __init__(self, optimizer, T_period, restarts=None, weights=None, eta_min=0, last_epoch=-1):
        self.optimizer = optimizer
        self.T_period = list(T_period)
        self.eta_min = eta_min if eta_min is not None else 0
        self.restarts = [v + 1 for v in (restarts if restarts else [0])]
        self.restart_weights = list(weights) if weights else [1]
        assert len(self.restarts) == len(self.restart_weights), 'restarts and their weights do not match.'
        for g in optimizer.param_groups:
            g.setdefault('initial_lr', g['lr'])
        self.last_epoch = last_epoch
        self.step()

    def _phase(self, t):
        """(iterations since the governing restart, its period, its weight)."""
        start, period, w = 0, self.T_period[0], 1.0
        for k, r in enumerate(self.restarts):
            if start <= r <= t:
                start, period, w = r, self.T_period[k + 1], self.restart_weights[k]     # IndexError as in the reference (:53)
        return t - start, period, w

    def get_lr(self):
        dt, period, w = self._phase(self.last_epoch)
        shape = (1.0 + math.cos(math.pi * dt / period)) / 2.0
        out = []
        for g in self.optimizer.param_groups:
            amp = g['initial_lr'] * (w if dt <= period else 1.0)
            out.append(self.eta_min + (amp - self.eta_min) * shape)
        return out

    def step(self, epoch=None):
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        for g, lr in zip(self.optimizer.param_groups, self.get_lr()):
            g['lr'] = lr

    def state_dict(self):
        return {'last_epoch': self.last_epoch}

    def load_state_dict(self, sd):
        self.last_epoch = sd['last_epoch']

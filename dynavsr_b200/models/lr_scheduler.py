"""Step schedule with restarts used by ``VideoBaseModel`` (``lr_scheme: MultiStepLR``; reference:
codes/models/lr_scheduler.py:8-32).  Closed form instead of torch's chained ``_LRScheduler`` recursion:

    lr(t) = initial_lr * w(last restart <= t) * gamma ** #{milestones m : last restart < m <= t}

with t counted in ``step()`` calls (the reference steps the scheduler once per iteration, base_model.py:51-53)."""


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

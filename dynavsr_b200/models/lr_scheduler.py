"""Learning-rate schedules with restarts used by ``VideoBaseModel`` (``lr_scheme: MultiStepLR`` /
``CosineAnnealingLR_Restart``; reference: codes/models/lr_scheduler.py:8-64).

The reference schedulers are torch ``_LRScheduler`` subclasses in *chained* form: every ``step()`` derives the new rate from
the group's CURRENT ``lr``, not from ``initial_lr``.  That is observable, and kept here: whatever else writes a group's rate
-- the linear warm-up of ``update_learning_rate`` (base_model.py:55-65), ``set_params_lr_zero`` of the ``ft_tsa_only``
option (Video_base_model.py:161-167) -- is carried forward by the schedule (a rate set to 0 stays 0 until the next restart;
after a warm-up the schedule continues from the last warm-up value).  Only a restart re-anchors at ``initial_lr * weight``.

    multi-step:  lr(t) = lr(t-1) * gamma ** #{milestones equal to t}
    cosine:      lr(t) = eta_min + (lr(t-1) - eta_min) * c(t) / c(t-1),   c(t) = 1 + cos(pi * (t - s) / T),
                 s = the last restart (0 before the first), T = the period that starts there; when c(t-1) = 0 (the step
                 after a trough) lr(t) = lr(t-1) + (initial_lr - eta_min) * (1 - cos(pi / T)) / 2   (:56-60)

t counts ``step()`` calls (one per iteration, base_model.py:51-53).  Both classes are checked against traces of the
unmodified reference classes, with and without outside writes to the rates (tests/golden/lr_schedules.json)."""
import math


class _RestartSchedule(object):
    def __init__(self, optimizer, restarts, weights, last_epoch):
        self.optimizer = optimizer
        self.restarts = [v + 1 for v in (restarts if restarts else [0])]
        self.restart_weights = list(weights) if weights else [1]
        assert len(self.restarts) == len(self.restart_weights), 'restarts and their weights do not match.'
        for g in optimizer.param_groups:
            g.setdefault('initial_lr', g['lr'])
        self.base_lrs = [g['initial_lr'] for g in optimizer.param_groups]
        self.last_epoch = last_epoch

    def _restart_weight(self, t):
        return self.restart_weights[self.restarts.index(t)] if t in self.restarts else None

    def step(self, epoch=None):
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        for g, lr in zip(self.optimizer.param_groups, self.get_lr()):
            g['lr'] = lr

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != 'optimizer'}

    def load_state_dict(self, sd):
        self.__dict__.update(sd)


class MultiStepLR_Restart(_RestartSchedule):
    def __init__(self, optimizer, milestones, restarts=None, weights=None, gamma=0.1, clear_state=False, last_epoch=-1):
        super(MultiStepLR_Restart, self).__init__(optimizer, restarts, weights, last_epoch)
        self.milestones = sorted(milestones or [])
        self.gamma = gamma
        self.clear_state = clear_state
        self.step()

    def get_lr(self):
        t = self.last_epoch
        w = self._restart_weight(t)
        if w is not None:
            if self.clear_state:
                self._clear_optimizer_state()
            return [g['initial_lr'] * w for g in self.optimizer.param_groups]
        decay = self.gamma ** self.milestones.count(t)
        return [g['lr'] * decay for g in self.optimizer.param_groups]

    def _clear_optimizer_state(self):
        """``clear_state`` (:25-26): forget the optimiser's moments at a restart."""
        flat = getattr(self.optimizer, 'flat', None)
        if flat is not None:
            if flat.m is not None:
                flat.m.zero_()
                flat.v.zero_()
        else:
            self.optimizer.state.clear()


class CosineAnnealingLR_Restart(_RestartSchedule):
    def __init__(self, optimizer, T_period, restarts=None, weights=None, eta_min=0, last_epoch=-1):
        super(CosineAnnealingLR_Restart, self).__init__(optimizer, restarts, weights, last_epoch)
        self.T_period = list(T_period)
        self.T_max = self.T_period[0]
        self.eta_min = eta_min if eta_min is not None else 0
        self.last_restart = 0
        self.step()

    def get_lr(self):
        t, eta = self.last_epoch, self.eta_min
        if t == 0:
            return list(self.base_lrs)
        w = self._restart_weight(t)
        if w is not None:
            self.last_restart = t
            self.T_max = self.T_period[self.restarts.index(t) + 1]          # IndexError when no period follows, as in :53
            return [g['initial_lr'] * w for g in self.optimizer.param_groups]
        dt, T = t - self.last_restart, self.T_max
        if (dt - 1 - T) % (2 * T) == 0:                                     # the previous step sat in a trough
            bump = (1.0 - math.cos(math.pi / T)) / 2.0
            return [g['lr'] + (base - eta) * bump for base, g in zip(self.base_lrs, self.optimizer.param_groups)]
        ratio = (1.0 + math.cos(math.pi * dt / T)) / (1.0 + math.cos(math.pi * (dt - 1) / T))
        return [ratio * (g['lr'] - eta) + eta for g in self.optimizer.param_groups]

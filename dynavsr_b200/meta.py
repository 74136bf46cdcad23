"""Outer (meta) step of DynaVSR training -- codes/train_dynavsr.py:252-438 -- on the CUDA path.

Per outer step every rank loops over its own tasks (clips; DistIterSampler sharding, dist.dist_iter_sampler_indices):

    theta' <- theta                                    (:322  deepcopy(model.netG), deepcopy(est_model.netE))
    K times:  SLR = MFDN'(LR);  L = w*cri(EDVR'(SLR), LR_c) + L1(SLR, SLR_true);  backward;  inner Adam/SGD on theta'
                                                       (:355-399; two lr groups lr_alpha / lr_alpha_est :335-344)
    L_q = w*cri(EDVR'(LR), HR_c);   meta_grad[G] += grad(L_q / B)            (:404-415)
    L_e = loss_e(MFDN'(LR), SLR_true);  meta_grad[E] += grad(L_e / (10 B))   (:417-426)
  one all-reduce (mean) of the FLAT meta-gradient over ranks, then ONE fused Adam/SGD launch on theta   (:438)

This is first-order MAML, the evident intent of the reference (its validation loop :633-677 and test_dynavsr.py:233-277
adapt the working copy exactly like this).  AS WRITTEN the training loop binds the inner optimiser to the copy but the
loss to the original (:322-399), so nothing is ever adapted and the inner losses' gradients leak, unscaled, into the
outer gradient; ``reference_quirk=True`` reproduces that accumulation (meta_grad = sum_tasks [K*grad L_inner(theta) +
grad L_q(theta)/B + grad L_e(theta)/(10B)]) for a numerical side-by-side on one GPU.

B200-first structure: theta' (EDVR u MFDN) is one flat buffer, so "deepcopy" = one D2D copy, the inner update = one launch,
``meta_grad += grad`` = one axpy over the flat gradient (kernels accumulate weight gradients straight into it).  The
exchange step and the outer update are ONE kernel over NVLink peer memory: the flat meta-gradient lives in symmetric memory
(torch.distributed._symmetric_memory); in the default reduce-scatter form (``dvsr_update_peers_sliced``) every rank reduces its
slice of all ranks' buffers in rank order, applies Adam / SGD to that slice with its slice of the moments and writes the new
weights into every rank's buffer (``exchange='peer-all'``: every rank reads all ranks' full gradients, ``dvsr_update_peers``) --
instead of DDP's bucketed hooks firing inside the inner loop, a reduced-gradient round trip through HBM and a separate optimiser
launch.  ``exchange='nccl'`` (or a process group without peer access, e.g. gloo in CPU tests) falls back to ONE NCCL all-reduce
of the flat buffer + the fused update launch.  ``MetaPool`` runs the tasks of an outer step on several task lanes (own working
copy, packs, CUDA graphs and stream each) side by side: a task is ~1 300 launches of small-patch kernels and leaves most of the
GPU idle on its own.
"""
import copy
import ctypes

import torch

from . import dist as ddist
from . import ops
from ._lib import call
from .adapt import FlatParams


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class MetaLearner(object):
    """netG: models.archs.EDVR_arch.EDVR, netE: models.archs.LRimg_estimator.DirectKernelEstimatorVideo (both on cuda).
    Their parameters become views of the working copy theta'; ``self.theta`` holds the meta-weights."""

    def __init__(self, netG, netE, inner_steps=1, lr_alpha=1e-5, lr_alpha_est=None, inner_optimizer='Adam',
                 inner_betas=(0.9, 0.99), criterion='cb', pixel_weight=1.0, est_loss='l1', outer_optimizer='Adam',
                 lr_outer=1e-5, outer_betas=(0.9, 0.99), reference_quirk=False, exchange='peer', use_graphs=False,
                 precision=None, policy=None):
        if inner_optimizer not in ('SGD', 'Adam') or outer_optimizer not in ('SGD', 'Adam'):
            raise NotImplementedError()
        if criterion not in ('l1', 'l2', 'cb', 'huber'):
            raise NotImplementedError('Loss type [%s] is not recognized.' % criterion)
        self.netG, self.netE = netG, netE
        self.K, self.lr_alpha = inner_steps, lr_alpha
        self.lr_alpha_est = lr_alpha if lr_alpha_est is None else lr_alpha_est
        self.inner_optimizer, self.inner_betas = inner_optimizer, inner_betas
        self.criterion, self.pixel_weight, self.est_loss = criterion, pixel_weight, est_loss
        self.outer_optimizer, self.lr_outer, self.outer_betas = outer_optimizer, lr_outer, outer_betas
        self.reference_quirk = reference_quirk
        # One CUDA graph per (task shape, tasks per step): the outer step of EDVR-L is ~4 000 launches of small-patch kernels and
        # was launch-bound when issued eagerly (47.9 ms, profiles/r1_meta_bench.txt).  The as-written reference mode keeps the
        # eager path (it exists for numerical side-by-sides only).
        self.use_graphs = bool(use_graphs) and not reference_quirk
        # operand precision of the tensor-core convolutions of a task: None = the backend's (bf16x3, fp32-class), 'bf16' = plain
        # bf16 operands with fp32 accumulation over the fp32 master weights -- BASELINE config 4's "bf16 compute, fp32 masters"
        self.precision = precision
        self._graphs = {}
        self.replayed_launches = 0      # library kernels launched through graph replays (bench.py's gpu_launches)
        self.exchange_events = []
        self.scope = ops.new_scope(policy)
        self.work = FlatParams([netG, netE], scope=self.scope)      # theta' (+ its gradient buffer)
        self.theta = self.work.meta                                  # theta: restored into theta' per task
        self.meta_grad = torch.zeros_like(self.theta)
        self._peer = None
        self.exchange = 'local'
        if exchange != 'none' and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            self.exchange = 'nccl'
            if exchange in ('peer', 'peer-all') and torch.distributed.get_backend() == 'nccl':
                try:        # flat meta-gradient in symmetric memory: peers read it directly (fused exchange + update)
                    import torch.distributed._symmetric_memory as symm
                    g = symm.empty(self.theta.numel(), dtype=torch.float32, device=self.theta.device)
                    hdl = symm.rendezvous(g, torch.distributed.group.WORLD)
                    g.zero_()
                    self.meta_grad, self._peer, self.exchange = g, hdl, exchange
                except Exception as e:      # no peer access / no symmetric-memory support in this build
                    import warnings
                    warnings.warn('MetaLearner: symmetric memory unavailable (%s); using NCCL all-reduce' % (e,))
        self.m = torch.zeros_like(self.theta) if outer_optimizer == 'Adam' else None
        self.v = torch.zeros_like(self.theta) if outer_optimizer == 'Adam' else None
        self.outer_steps = 0
        self.N, self.center = netG.nframes, netG.center
        self.last = {}

    def _eps(self):
        return 1e-2 if self.criterion == 'huber' else 1e-6

    # ------------------------------------------------------------------ one task
    def _task(self, task, n_tasks):
        if not self.use_graphs:
            with ops.conv_precision(self.precision):
                return self._task_eager(task, n_tasks)
        key = tuple(tuple(task[k].shape) for k in ('LQs', 'GT', 'SuperLQs')) + (n_tasks,)
        st = self._graphs.get(key) or self._capture_task(key, task, n_tasks)
        for k in ('LQs', 'GT', 'SuperLQs'):
            st['in'][k].copy_(task[k], non_blocking=True)
        st['graph'].replay()
        self.replayed_launches += st['launches']
        lq, le, inner = st['out']                   # static tensors of the graph: copy out before the next replay overwrites them
        return lq.clone(), le.clone(), [t.clone() for t in inner]

    def _capture_task(self, key, task, n_tasks):
        """Warm up and capture one task (theta' <- theta, K inner steps, meta-test backward, meta_grad += grad) as a CUDA graph
        over static input buffers.  The packs of theta are restored from the scope's snapshot arena inside the graph; outer_step
        refreshes that snapshot after every meta-update."""
        fl = self.work
        st = {'in': {k: task[k].clone() for k in ('LQs', 'GT', 'SuperLQs')}}
        keep = self.meta_grad.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), ops.conv_precision(self.precision):
            for _ in range(2):                      # allocator / autograd / pack-registry warm-up
                self._task_eager(st['in'], n_tasks)
        torch.cuda.current_stream().wait_stream(side)
        self.meta_grad.copy_(keep)
        fl.restore()
        ops.repack_all()
        ops.snapshot_packs()
        g = torch.cuda.CUDAGraph()
        from . import _lib
        n0 = _lib.COUNTER[0]
        with torch.cuda.graph(g), ops.conv_precision(self.precision):
            st['out'] = self._task_eager(st['in'], n_tasks, from_snapshot=True)
        st['launches'] = _lib.COUNTER[0] - n0        # library kernels inside one replay of this graph
        st['graph'] = g
        self._graphs[key] = st
        self.meta_grad.copy_(keep)                  # capture does not execute, but keep the invariant explicit
        return st

    def _task_eager(self, task, n_tasks, from_snapshot=False):
        """task: dict of device tensors  LQs [1, N, 3, h, w], GT [1, N, 3, sh, sw] or [1, 3, sh, sw], SuperLQs [1, N, 3, h/s, w/s]."""
        LQs, GT, SLQ = task['LQs'], task['GT'], task['SuperLQs']
        B, N, C, h, w = LQs.shape
        gt_hr = GT[:, self.center] if GT.dim() == 5 else GT
        frames = ops.to_nhwc(LQs.reshape(B * N, C, h, w))
        lr_c = frames.view(B, N, h, w, C)[:, self.center].contiguous()             # meta_train GT = LR centre (:292)
        slr_true = ops.to_nhwc(SLQ.reshape(B * N, C, SLQ.shape[-2], SLQ.shape[-1]))
        hr_c = ops.to_nhwc(gt_hr)
        fl = self.work
        if not self.reference_quirk:
            fl.restore()                                                            # theta' <- theta (:322)
            if from_snapshot:
                if not ops.restore_packs():                                         # one copy launch from the snapshot arena
                    raise RuntimeError('MetaLearner: weight-pack snapshot invalid during CUDA-graph capture')
            else:
                ops.repack_all()
        inner = []
        for k in range(self.K):
            if not self.reference_quirk:
                fl.zero_grad()
            slr = self.netE.forward_nhwc(frames, B, N)                              # :360-363
            sr = self.netG.forward_nhwc(slr, B, N)                                  # :384-385
            loss = ops.pixel_loss(sr, lr_c, self.criterion, self.pixel_weight, self._eps()) + \
                ops.pixel_loss(slr, slr_true, 'l1', 1.0)                            # :395
            loss.backward()                                                         # :397
            inner.append(loss.detach())
            if not self.reference_quirk:                                            # :399 (a no-op as written)
                if self.inner_optimizer == 'SGD':
                    fl.sgd_step(self.lr_alpha, self.lr_alpha_est)
                else:
                    fl.adam_step(self.lr_alpha, self.lr_alpha_est, self.inner_betas, step=k + 1)
        if not self.reference_quirk:
            fl.zero_grad()
        # meta test at theta' (:404-426); the two losses touch disjoint halves of the flat gradient
        sr = self.netG.forward_nhwc(frames, B, N)
        loss_q = ops.pixel_loss(sr, hr_c, self.criterion, self.pixel_weight, self._eps())
        (loss_q / n_tasks).backward()
        slr = self.netE.forward_nhwc(frames, B, N)
        loss_e = ops.pixel_loss(slr, slr_true, self.est_loss, 1.0)
        (loss_e / (n_tasks * 10)).backward()
        ops.join_async(self.scope)
        if not self.reference_quirk:
            self.meta_grad.add_(fl.grad)
        return loss_q.detach(), loss_e.detach(), inner

    # ------------------------------------------------------------------ one outer step
    def outer_step(self, tasks, lr=None):
        """tasks: this rank's clips for the step.  Returns the summed query loss / B (the reference's ``total_loss_q``)."""
        with ops.scope(self.scope):
            lq, le, inner = self._run_tasks(tasks, len(tasks))
            return self._finish_step(lq, le, inner, len(tasks), lr)

    def _run_tasks(self, tasks, n_tasks):
        """Zero this learner's meta-gradient and accumulate the meta-gradients of ``tasks`` into it (each scaled for a step of
        ``n_tasks`` tasks).  Call inside ``ops.scope(self.scope)``."""
        fl = self.work
        self.meta_grad.zero_()
        if self.reference_quirk:
            fl.restore()
            ops.repack_all()
            fl.zero_grad()                     # optimizer.zero_grad() (:270): grads then pile up across tasks
        if fl.m is not None:
            fl.m.zero_(); fl.v.zero_()
        lq, le, inner = [], [], []
        for t in tasks:
            if self.inner_optimizer == 'Adam' and fl.m is not None:
                fl.m.zero_(); fl.v.zero_()     # a fresh inner optimiser per task (:346-353)
            a, b, c = self._task(t, n_tasks)
            lq.append(a); le.append(b); inner.append(c)
        if self.reference_quirk:
            self.meta_grad.copy_(fl.grad)
        return lq, le, inner

    def _finish_step(self, lq, le, inner, n_tasks, lr=None):
        """Exchange of the flat meta-gradient + outer update (one kernel on the peer paths), then the modules and packs are
        brought to the new meta-weights.  Call inside ``ops.scope(self.scope)``."""
        fl = self.work
        if True:
            lr = self.lr_outer if lr is None else lr
            self.outer_steps += 1
            n = self.theta.numel()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            if self._peer is not None:
                # ---- exchange + outer update in ONE kernel over NVLink peer memory
                h = self._peer
                world = torch.distributed.get_world_size()
                b1, b2 = self.outer_betas
                t = self.outer_steps
                adam = self.outer_optimizer == 'Adam'
                h.barrier(channel=0)               # every rank's meta-gradient is complete
                if self.exchange == 'peer':
                    # reduce-scatter form (update.cu): this rank reduces + updates its slice and writes the new weights of the slice
                    # into every rank's exchange buffer; after the barrier the buffer IS the new theta (moments are sharded by slice)
                    call('dvsr_update_peers_sliced', _p(self.theta), ctypes.c_void_p(int(h.buffer_ptrs_dev)), world,
                         torch.distributed.get_rank(), int(getattr(h, 'offset', 0)) // 4, 1.0 / world, _p(self.m), _p(self.v), n, n,
                         float(lr), float(lr), float(b1), float(b2), 1e-8, float(1 - b1 ** t) if adam else 1.0,
                         float(1 - b2 ** t) if adam else 1.0, 0.0, 1 if adam else 0, _stream())
                    h.barrier(channel=1)           # every slice of every buffer has been written
                    self.theta.copy_(self.meta_grad)
                else:
                    # 'peer-all': every rank reads all ranks' full gradients and updates its whole copy (fewest barriers' worth of
                    # logic, fine for 2 ranks; traffic grows with the rank count)
                    call('dvsr_update_peers', _p(self.theta), ctypes.c_void_p(int(h.buffer_ptrs_dev)), world,
                         int(getattr(h, 'offset', 0)) // 4, 1.0 / world, _p(self.m), _p(self.v), n, n, float(lr), float(lr), float(b1),
                         float(b2), 1e-8, float(1 - b1 ** t) if adam else 1.0, float(1 - b2 ** t) if adam else 1.0, 0.0,
                         1 if adam else 0, _stream())
                    h.barrier(channel=1)           # nobody overwrites its gradient while a peer still reads it
            elif self.exchange == 'nccl':
                # ---- fallback: ONE all-reduce (mean) of the flat meta-gradient, then the fused update launch
                ddist.allreduce_flat_gradient(self.meta_grad, average=True)
            if self._peer is not None:
                pass
            elif self.outer_optimizer == 'SGD':
                call('dvsr_update_sgd', _p(self.theta), _p(self.meta_grad), n, n, float(lr), float(lr), 0.0, _stream())
            else:
                b1, b2 = self.outer_betas
                t = self.outer_steps
                call('dvsr_update_adam', _p(self.theta), _p(self.meta_grad), _p(self.m), _p(self.v), n, n, float(lr), float(lr),
                     float(b1), float(b2), 1e-8, float(1 - b1 ** t), float(1 - b2 ** t), 0.0, _stream())
            ev1.record()
            if len(self.exchange_events) < 256:
                self.exchange_events.append((ev0, ev1))
            ops.invalidate_pack_snapshot(self.scope)
            fl.restore()                           # the modules now hold the updated meta-weights
            ops.repack_all()
            if self.use_graphs:
                ops.snapshot_packs()               # the captured tasks restore theta's packs from this arena
        self.last = {'loss_q': torch.stack(lq), 'loss_e': torch.stack(le), 'inner': inner}
        return torch.stack(lq).sum() / n_tasks

    def exchange_timing(self, reset=True):
        """Device time of the exchange + outer-update section of the recorded outer steps (CUDA events on the launching stream;
        includes the two symmetric-memory barriers of the fused path) and the NVLink read rate it implies for the fused kernel:
        every rank reads (world - 1) peers' flat gradients."""
        if not self.exchange_events:
            return None
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in self.exchange_events)
        if reset:
            self.exchange_events = []
        world = torch.distributed.get_world_size() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        med = ms[len(ms) // 2]
        nbytes = self.theta.numel() * 4
        # NVLink bytes per rank: 'peer' (sliced) reads (world-1)/world of the buffer and writes the same amount; 'peer-all' reads
        # (world-1) whole buffers; NCCL's ring / tree moves 2 (world-1)/world of it
        peer_read = (world - 1) * nbytes // world if self.exchange in ('peer', 'nccl') else (world - 1) * nbytes
        peer_write = (world - 1) * nbytes // world if self.exchange in ('peer', 'nccl') else 0
        return {'path': self.exchange, 'median_us': med * 1e3, 'min_us': ms[0] * 1e3, 'flat_gradient_bytes': nbytes,
                'peer_bytes_read_per_rank': peer_read, 'peer_bytes_written_per_rank': peer_write,
                'nvlink_GBps_per_rank': (max(peer_read, peer_write) / (med * 1e-3) / 1e9) if world > 1 else None,
                'local_hbm_GBps': ((3 if self.outer_optimizer == 'SGD' else 7) * nbytes / (med * 1e-3) / 1e9)}

    def dtype_string(self):
        if not ops._tc():
            return 'f32 (exact CUDA-core path)'
        prec = self.precision or ops._backend['precision']
        if getattr(self.netG, 'nf', 64) > 64:
            # 128-channel layers exceed the resident-weight kernel's shared-memory budget and run on the streaming tcgen05 kernel
            # (TF32 operands rounded to nearest); only the <= 64-channel layers (MFDN, head) follow `precision`
            return ('f32 storage, fp32 master weights / gradients / optimiser state; tcgen05 TF32 operands (streaming kernel, nf = %d '
                    'layers) and %s operands (<= 64-channel layers), fp32 accumulate' % (self.netG.nf, prec))
        return 'f32 storage, fp32 master weights / gradients / optimiser state; tcgen05 %s operands, fp32 accumulate' % prec

    def state_dicts(self):
        """(EDVR state_dict, MFDN state_dict) of the meta-weights, reference key names (checkpoint contract)."""
        return ({k: v.detach().clone() for k, v in self.netG.state_dict().items()},
                {k: v.detach().clone() for k, v in self.netE.state_dict().items()})


class MetaPool(object):
    """The tasks of an outer step on L task lanes of ONE GPU.

    A task of the meta step (train_dynavsr.py:322-426: theta' <- theta, K inner steps on the 16x16 SLR patch, the meta-test
    backward on the 64x64 LR patch) is a dependent chain of ~1 300 small-patch launches that keeps a few dozen SMs busy at a
    time; the tasks of a step are independent (they all start from theta and only ADD into the meta-gradient).  Each lane is a
    ``MetaLearner`` of its own -- working copy theta', flat gradient, weight packs, CUDA graphs, stream -- so the lanes' tasks
    run side by side; their meta-gradients are summed into lane 0's, whose exchange + outer update (one kernel over NVLink peer
    memory) then runs once, and the new theta is copied back to the other lanes.  Numerically the step equals
    ``MetaLearner.outer_step`` on the same tasks up to the fp32 order of the cross-lane sum.
    """

    def __init__(self, netG, netE, lanes=4, **kw):
        assert lanes >= 1 and not kw.get('reference_quirk'), 'the as-written reference mode is sequential by construction'
        clones = [(copy.deepcopy(netG), copy.deepcopy(netE)) for _ in range(lanes - 1)]     # BEFORE lane 0 re-homes the parameters
        lane_kw = dict(kw)
        lane_kw['exchange'] = 'none'
        self.master = MetaLearner(netG, netE, **kw)
        self.lanes = [self.master] + [MetaLearner(g, e, **lane_kw) for g, e in clones]
        self.streams = [torch.cuda.Stream() for _ in self.lanes]
        self._warm = set()

    # the attributes callers of MetaLearner read
    theta = property(lambda self: self.master.theta)
    meta_grad = property(lambda self: self.master.meta_grad)
    exchange = property(lambda self: self.master.exchange)
    use_graphs = property(lambda self: self.master.use_graphs)
    last = property(lambda self: self.master.last)

    @property
    def replayed_launches(self):
        return sum(l.replayed_launches for l in self.lanes)

    @replayed_launches.setter
    def replayed_launches(self, v):
        for l in self.lanes:
            l.replayed_launches = v

    @property
    def exchange_events(self):
        return self.master.exchange_events

    @exchange_events.setter
    def exchange_events(self, v):
        self.master.exchange_events = v

    def exchange_timing(self, reset=True):
        return self.master.exchange_timing(reset)

    def dtype_string(self):
        return self.master.dtype_string()

    def state_dicts(self):
        return self.master.state_dicts()

    def outer_step(self, tasks, lr=None):
        m, L = self.master, len(self.lanes)
        n_tasks = len(tasks)
        cur = torch.cuda.current_stream()
        share = [list(range(li, n_tasks, L)) for li in range(L)]
        out = [None] * L
        # a lane's first task of a given shape captures its CUDA graph: do that with the GPU to itself (capture is a global mode)
        key = tuple(tuple(tasks[0][k].shape) for k in ('LQs', 'GT', 'SuperLQs')) + (n_tasks,)
        serial = key not in self._warm
        for li, lane in enumerate(self.lanes):
            if not share[li]:
                continue
            s = self.streams[li]
            s.wait_stream(cur)
            with torch.cuda.stream(s), ops.scope(lane.scope):
                out[li] = lane._run_tasks([tasks[i] for i in share[li]], n_tasks)
            if serial:
                s.synchronize()
        self._warm.add(key)
        for li, s in enumerate(self.streams):
            if share[li]:
                cur.wait_stream(s)
        lq, le, inner = [None] * n_tasks, [None] * n_tasks, [None] * n_tasks
        for li in range(L):
            if share[li]:
                for j, i in enumerate(share[li]):
                    lq[i], le[i], inner[i] = out[li][0][j], out[li][1][j], out[li][2][j]
                if li > 0:
                    m.meta_grad.add_(self.lanes[li].meta_grad)          # cross-lane sum, on the caller's stream
        with ops.scope(m.scope):
            loss = m._finish_step(lq, le, inner, n_tasks, lr)
        for lane in self.lanes[1:]:                                     # the other lanes follow the new meta-weights
            lane.theta.copy_(m.theta)
            with ops.scope(lane.scope):
                ops.invalidate_pack_snapshot(lane.scope)
                lane.work.restore()
                ops.repack_all()
                if lane.use_graphs:
                    ops.snapshot_packs()
        return loss

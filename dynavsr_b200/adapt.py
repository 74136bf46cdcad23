"""Test-time adaptation engine: the hot loop of codes/test_dynavsr.py:197-283 for one output frame.

    restore meta-weights                      (:208  deepcopy(model.netG), deepcopy(est_model.netE))
    repeat adapt_iter times:                  (:235)
        SLR      = MFDN_cp(LR)                (:238-241)
        loss     = cri(EDVR_cp(SLR), LR_c)    (:262-264)
                 + 10 * L1(SLR, MFDN_fixed(LR))   (:267-274; MFDN_fixed(LR) is constant -> hoisted)
        backward; SGD / Adam step on EDVR+MFDN    (:276-277)
    HR = EDVR_cp(LR)                          (:282-283)

B200-first structure: EDVR and MFDN parameters live in ONE flat fp32 buffer (and one flat gradient
buffer), so "deepcopy" is a single device-to-device copy, ``zero_grad`` a single memset and the
optimiser step a single kernel launch (dvsr_update_sgd / dvsr_update_adam) instead of 158 per-tensor
launches; weight-gradient kernels accumulate straight into the flat gradient buffer.  The whole
step (forward, backward, update) and the final forward are captured in CUDA graphs per input shape.
"""
import copy
import ctypes
import weakref

import torch

from . import _lib, ops
from ._lib import call


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class FlatParams(object):
    """Re-homes the parameters of ``modules`` (in order) into one flat buffer + one flat gradient buffer.

    ``split`` is the element offset where the second module's parameters start (the two learning-rate
    groups of train_dynavsr.py:335-344: lr_alpha for EDVR, lr_alpha_est for MFDN).
    """
    ALIGN = 64  # floats (256 B)

    def __init__(self, modules, device=None, scope=None):
        """``modules``: list of nn.Modules, or list of parameter lists (explicit groups, e.g. the optimiser param groups
        of Video_base_model.py:57-126); the second entry starts the second learning-rate group."""
        self.scope = scope or ops._default_scope
        params = []
        self.split = None
        offsets = []
        n = 0
        for mi, m in enumerate(modules):
            if mi == 1:
                self.split = n
            for p in (m.parameters() if hasattr(m, 'parameters') else m):
                offsets.append(n)
                params.append(p)
                n += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if self.split is None:
            self.split = n
        for p in params:
            owner = getattr(p, '_dvsr_flat', None)
            owner = owner() if owner is not None else None
            if owner is not None and owner is not self:
                # a parameter lives in exactly one flat buffer: re-homing it again would orphan the first owner's buffers
                # (its optimiser would then step a buffer no parameter points to).  Release the previous owner explicitly.
                raise RuntimeError('parameter already belongs to another FlatParams (e.g. the FlatOptimizer create_model built '
                                   'for is_train); call its .release() first, or hand that FlatParams to the new engine')
        device = device or params[0].device
        self.numel = n
        self.flat = torch.zeros(n, device=device, dtype=torch.float32)
        self.grad = torch.zeros(n, device=device, dtype=torch.float32)
        self.params, self.offsets = params, offsets
        with torch.no_grad():
            for p, o in zip(params, offsets):
                view = self.flat[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = None
                p._dvsr_grad = self.grad[o:o + p.numel()].view(p.shape)   # kernels accumulate here directly
                p._dvsr_scope = self.scope
                p._dvsr_flat = weakref.ref(self)
        self.released = False
        self.meta = self.flat.clone()       # the meta-weights every frame restarts from
        self.m = self.v = None
        self.step_count = 0
        ops.invalidate_weight_cache(self.scope)

    def release(self):
        """Give the parameters back: each gets its own storage again (current values), the flat-gradient slots and the scope
        tag are removed, and every later use of this object raises.  After this another FlatParams / engine may adopt them."""
        with torch.no_grad():
            for p in self.params:
                p.data = p.data.clone()
                for a in ('_dvsr_grad', '_dvsr_scope', '_dvsr_flat'):
                    if hasattr(p, a):
                        delattr(p, a)
        self.released = True
        ops.invalidate_weight_cache(self.scope)

    def _alive(self):
        if self.released:
            raise RuntimeError('this FlatParams was released: its parameters now live elsewhere')

    def snapshot(self):
        self._alive()
        self.meta.copy_(self.flat)
        ops.invalidate_pack_snapshot(self.scope)

    def restore(self):
        """test_dynavsr.py:208 -- one D2D copy instead of two module deep-copies."""
        self._alive()
        self.flat.copy_(self.meta)
        self.step_count = 0
        if self.m is not None:
            self.m.zero_()
            self.v.zero_()
        ops.invalidate_weight_cache(self.scope)      # the step that follows starts with ops.repack_all()

    def zero_grad(self):
        self.grad.zero_()

    def fold_autograd_grads(self):
        """Gradients that reached a parameter through ``p.grad`` (ops that are not this library's: a torch loss on the weights,
        a regulariser) are added to the flat gradient and cleared, so the fused update never silently ignores them."""
        for p in self.params:
            if p.grad is not None:
                p._dvsr_grad.add_(p.grad)
                p.grad = None

    def sgd_step(self, lr0, lr1, weight_decay=0.0):
        self._alive()
        self.fold_autograd_grads()
        ops.join_async(self.scope)        # weight-gradient kernels run on this scope's side stream
        call('dvsr_update_sgd', _p(self.flat), _p(self.grad), self.numel, self.split, float(lr0), float(lr1),
             float(weight_decay), _stream())
        ops.weights_updated(self.scope)

    def adam_step(self, lr0, lr1, betas=(0.9, 0.999), eps=1e-8, step=None, weight_decay=0.0):
        self._alive()
        self.fold_autograd_grads()
        if self.m is None:
            self.m = torch.zeros_like(self.flat)
            self.v = torch.zeros_like(self.flat)
        ops.join_async(self.scope)
        self.step_count = self.step_count + 1 if step is None else step
        t = self.step_count
        bc1, bc2 = 1.0 - betas[0] ** t, 1.0 - betas[1] ** t
        call('dvsr_update_adam', _p(self.flat), _p(self.grad), _p(self.m), _p(self.v), self.numel, self.split,
             float(lr0), float(lr1), float(betas[0]), float(betas[1]), float(eps), float(bc1), float(bc2),
             float(weight_decay), _stream())
        ops.weights_updated(self.scope)


class InnerLoopAdapter(object):
    """Resident adapt-then-super-resolve engine for one (netG, netE, netE_fixed) triple.

    netG : models.archs.EDVR_arch.EDVR                    (adapted copy; meta-weights = state at construction)
    netE : models.archs.LRimg_estimator.DirectKernelEstimatorVideo (adapted copy)
    netE_fixed : same class, frozen (the `fixed_E` checkpoint, test_dynavsr.py:211)
    """

    def __init__(self, netG, netE, netE_fixed, steps=2, lr_alpha=1e-5, lr_alpha_est=None, optimizer='SGD',
                 betas=(0.9, 0.99), criterion='l2', slr_weight=10.0, pixel_weight=1.0, use_graphs=True, inner_precision=None, policy=None,
                 use_real=False, use_patch=False, num_patch=1, patch_size=128):
        if optimizer not in ('SGD', 'Adam'):
            raise NotImplementedError(optimizer)
        if criterion not in ('l1', 'l2', 'cb'):
            raise NotImplementedError('Loss type [%s] is not recognized.' % criterion)
        self.netG, self.netE, self.netE_fixed = netG, netE, netE_fixed
        self.steps, self.optimizer, self.betas, self.criterion = steps, optimizer, betas, criterion
        self.lr_alpha = lr_alpha
        self.lr_alpha_est = lr_alpha if lr_alpha_est is None else lr_alpha_est
        self.slr_weight, self.pixel_weight = slr_weight, pixel_weight
        # the two optional branches of the reference loop: train.use_real (the dataset's pre-generated super-LR clip replaces
        # MFDN(LR) and only EDVR adapts, test_dynavsr.py:218-221,243-244) and train.maml.use_patch (the pixel loss is taken on
        # num_patch random crops, :118-145,255-260).  Crop positions are drawn per frame, so use_patch runs eagerly (no graph).
        self.use_real, self.use_patch, self.num_patch, self.patch_size = bool(use_real), bool(use_patch), int(num_patch), int(patch_size)
        self.use_graphs = use_graphs and not self.use_patch
        # operand precision of the tensor-core convolutions DURING the adaptation steps: None = the backend's setting, a name
        # ('bf16' = single product) or a (forward, backward) pair.  The steps' errors reach the frame attenuated by how little
        # the adaptation moves it (profiles/r1_precision_study.md: SGD tolerates 'bf16' throughout; Adam's normalised step does
        # not -- keep its forward at 'bf16x3', i.e. (None, 'bf16')).  The final forward always runs at the backend's precision.
        self.inner_precision = inner_precision
        # this engine's weight packs / pack table / weight-gradient side stream / launch policy of the persistent kernels
        self.scope = ops.new_scope(policy)
        self.flat = FlatParams([netG, netE], scope=self.scope)
        for p in netE_fixed.parameters():
            p.requires_grad_(False)
            p._dvsr_scope = self.scope
        self.N = netG.nframes
        self.center = netG.center
        self._graphs = {}
        self.last_losses = None
        self.launches_per_step = None

    # ------------------------------------------------------------------ eager building blocks
    def _draw_patches(self, h, w):
        """Crop positions of one inner step, drawn exactly as the reference does (preprocessing.common_crop :76-77 through
        test_dynavsr.crop :138-141: per patch ``py`` then ``px`` from Python's ``random``), in SLR pixels."""
        import random
        q = self.patch_size // 2
        if q > h or q > w:
            raise RuntimeError('use_patch: patch_size // 2 = %d exceeds the super-LR frame %dx%d' % (q, h, w))
        return [(random.randrange(0, h - q + 1), random.randrange(0, w - q + 1)) for _ in range(self.num_patch)]

    def _inner_step(self, frames, gt, slr_fixed, B, step_idx, slr_given=None, positions=None):
        """One adaptation step on channels-last tensors; returns the (device) loss scalar."""
        self.flat.zero_grad()
        p_fwd, p_bwd = self.inner_precision if isinstance(self.inner_precision, (tuple, list)) else (self.inner_precision,) * 2
        with ops.conv_precision(p_fwd):
            slr = slr_given if self.use_real else self.netE.forward_nhwc(frames, B, self.N)
            if self.use_patch:
                if B != 1:
                    raise RuntimeError('use_patch crops one clip into a batch of patches: B must be 1 (test_dynavsr.py:134)')
                q, s = self.patch_size // 2, self.netG.scale
                pos = positions if positions is not None else self._draw_patches(slr.shape[1], slr.shape[2])
                s_in = torch.cat([slr[:, py:py + q, px:px + q, :] for py, px in pos], 0).contiguous()            # [P*N, q, q, 3]
                s_gt = torch.cat([gt[:, s * py:s * (py + q), s * px:s * (px + q), :] for py, px in pos], 0).contiguous()
                sr, target = self.netG.forward_nhwc(s_in, len(pos), self.N), s_gt
            else:
                sr, target = self.netG.forward_nhwc(slr, B, self.N), gt
            loss = ops.pixel_loss(sr, target, self.criterion, self.pixel_weight) + \
                ops.pixel_loss(slr, slr_fixed, 'l1', self.slr_weight)
        with ops.conv_precision(p_bwd):
            loss.backward()
        if self.optimizer == 'SGD':
            self.flat.sgd_step(self.lr_alpha, self.lr_alpha_est)
        else:
            self.flat.adam_step(self.lr_alpha, self.lr_alpha_est, self.betas, step=step_idx + 1)
        return loss.detach()

    def _restore_packs(self):
        """The weights were restored to the meta-weights just before: bring the kernel-layout packs back too -- a copy
        of their snapshot when there is one, else one table-driven re-pack launch (whose result becomes the snapshot)."""
        if not ops.restore_packs():
            ops.repack_all()
            if not torch.cuda.is_current_stream_capturing():
                ops.snapshot_packs()

    def _run_eager(self, frames, B, slr_given=None, positions=None):
        H, W = frames.shape[1], frames.shape[2]
        self._restore_packs()
        gt = frames.view(B, self.N, H, W, 3)[:, self.center].contiguous()
        with torch.no_grad():
            slr_fixed = self.netE_fixed.forward_nhwc(frames, B, self.N)
        losses = [self._inner_step(frames, gt, slr_fixed, B, i, slr_given, positions[i] if positions is not None else None)
                  for i in range(self.steps)]
        with torch.no_grad():
            hr = self.netG.forward_nhwc(frames, B, self.N)
        return hr, losses

    # ------------------------------------------------------------------ CUDA-graph path
    def _build_graphs(self, frames, B, slr=None):
        key = (tuple(frames.shape), B)
        st = {'in': frames.clone(), 'slr': slr.clone() if slr is not None else None}
        # warm-up on a side stream (allocator + autograd warm-up, as PyTorch's capture recipe requires)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.flat.restore()
            self._run_eager(st['in'], B, st['slr'])
            self.flat.restore()
        torch.cuda.current_stream().wait_stream(side)
        ops.repack_all()            # (re)build the device-side pack table now: it must not change during capture
        ops.snapshot_packs()        # packs of the meta-weights: every replay restores them with one copy
        g = torch.cuda.CUDAGraph()
        n0 = _lib.COUNTER[0]
        with torch.cuda.graph(g):
            hr, losses = self._run_eager(st['in'], B, st['slr'])
            st['hr'], st['losses'] = hr, torch.stack(losses) if losses else None
        self.launches_per_step = _lib.COUNTER[0] - n0      # kernels of libdvsr_b200.so inside one replay
        st['graph'] = g
        self._graphs[key] = st
        return st

    # ------------------------------------------------------------------ public API
    def adapt_and_infer_nhwc(self, frames, B=1, slr=None, patch_positions=None):
        """frames: [B*N, H, W, 3] device tensor (LR window, channels-last) -> HR [B, sH, sW, 3].
        ``slr`` ([B*N, H/s, W/s, 3]) is the pre-generated super-LR clip of the ``use_real`` branch; ``patch_positions`` (per inner
        step a list of (py, px) in SLR pixels) overrides the random crops of the ``use_patch`` branch."""
        H, W = frames.shape[1], frames.shape[2]
        s = self.netG.scale
        if self.use_real and slr is None:
            raise RuntimeError("use_real: pass the dataset's super-LR clip (val_data['SuperLQs'], test_dynavsr.py:243-244) as slr=")
        if not self.use_real:
            slr = None
        if H % (4 * s) or W % (4 * s):
            raise RuntimeError('adaptation runs EDVR on LR/scale: H and W must be multiples of %d, got %dx%d '
                               '(the reference dataset crops for this: video_test_dataset_int.py:185-189)' % (4 * s, H, W))
        with ops.scope(self.scope):
            if not self.use_graphs:
                self.flat.restore()
                hr, losses = self._run_eager(frames, B, slr, patch_positions)
                self.last_losses = torch.stack(losses) if losses else None
                return hr
            st = self._graphs.get((tuple(frames.shape), B)) or self._build_graphs(frames, B, slr)
            if not self.scope.arena_valid:
                # the meta-weights changed since the graphs were captured (FlatParams.snapshot): refresh the pack
                # snapshot the captured restore-copy reads, in place
                self.flat.restore()
                ops.repack_all()
                ops.snapshot_packs()
            st['in'].copy_(frames, non_blocking=True)
            if slr is not None:
                st['slr'].copy_(slr, non_blocking=True)
            self.flat.restore()
            st['graph'].replay()
            self.last_losses = st['losses']
            return st['hr']

    def set_meta_weights(self, state_dict_G=None, state_dict_E=None, strict=True):
        """Replace the meta-weights every frame restarts from (e.g. after a meta-training step or when another checkpoint is
        loaded).  The working copy still holds the LAST frame's adapted weights, so it is first restored, then overwritten,
        then snapshotted; captured CUDA graphs pick the new weights (and their packs) up on the next call."""
        with ops.scope(self.scope):
            self.flat.restore()
            if state_dict_G is not None:
                self.netG.load_state_dict(state_dict_G, strict=strict)
            if state_dict_E is not None:
                self.netE.load_state_dict(state_dict_E, strict=strict)
            self.flat.snapshot()

    def adapt_and_infer(self, lr_clip, slr_clip=None, patch_positions=None):
        """lr_clip: [B, N, 3, H, W] (reference tensor layout, host or device) -> HR [B, 3, sH, sW]; slr_clip: [B, N, 3, H/s, W/s]
        (``val_data['SuperLQs']``) for the use_real branch."""
        B, N, C, H, W = lr_clip.shape
        x = lr_clip.to('cuda', non_blocking=True).reshape(B * N, C, H, W)
        frames = ops.to_nhwc(x)
        slr = None
        if slr_clip is not None:
            slr = ops.to_nhwc(slr_clip.to('cuda', non_blocking=True).reshape(B * N, C, slr_clip.shape[-2], slr_clip.shape[-1]))
        return ops.to_nchw(self.adapt_and_infer_nhwc(frames, B, slr, patch_positions))

    def infer_nhwc(self, frames, B=1):
        """Plain EDVR forward with the meta-weights (no adaptation)."""
        with ops.scope(self.scope):
            self.flat.restore()
            self._restore_packs()
            with torch.no_grad():
                return self.netG.forward_nhwc(frames, B, self.N)


class AdaptationPool(object):
    """P independent ``InnerLoopAdapter`` pipelines on P CUDA streams of ONE GPU.

    Output frames are independent units of work: every frame restarts from the same meta-weights and reads only its own
    window (test_dynavsr.py:208), which is why the path shards over GPUs with no collective (train_dynavsr.py:509).  The
    same property lets one GPU keep several frames in flight: the inner adaptation steps run EDVR on the 4x smaller SLR
    window (44x80 -> 28-140 CTAs per kernel, a dependent chain of ~400 launches), which leaves most of the 148 SMs idle;
    a second and third frame's kernels fill them.  Each pipeline owns its parameter copy (flat buffer), weight packs,
    CUDA graphs and streams; nothing is shared but the (read-only) meta-weights they were cloned from.
    """

    def __init__(self, netG, netE, netE_fixed, pipelines=6, cta_budget=None, min_tiles_per_cta=None, min_chunks_per_cta=None, **kw):
        assert pipelines >= 1
        # Launch policy of the persistent kernels while several frames share the GPU (measured on B200, bench.py sweep):
        # a launch is held to ~1/4 of the SMs and every conv CTA takes >= 2 tiles, so the pipelines' launches run side by
        # side instead of queueing behind each other's 148-CTA grids.  One pipeline keeps the whole GPU per launch.
        # The policy lives in each engine's PackScope and travels inside every descriptor (dvsr_policy): two pools with
        # different policies in one process do not interact.
        sms = _lib.lib().dvsr_sm_count() if torch.cuda.is_available() else 148
        self.cta_budget = cta_budget if cta_budget is not None else (sms if pipelines == 1 else (sms // 2 if pipelines < 4 else sms // 4))
        self.min_tiles_per_cta = min_tiles_per_cta if min_tiles_per_cta is not None else (1 if pipelines == 1 else 2)
        self.min_chunks_per_cta = min_chunks_per_cta if min_chunks_per_cta is not None else (4 if pipelines == 1 else 24)   # fewer split-K partial sums
        nets = [(netG, netE, netE_fixed)]
        for _ in range(pipelines - 1):          # clone BEFORE any engine re-homes the parameters
            nets.append((copy.deepcopy(netG), copy.deepcopy(netE), copy.deepcopy(netE_fixed)))
        self.engines = [InnerLoopAdapter(g, e, f, policy=ops.LaunchPolicy(self.cta_budget, self.min_tiles_per_cta, self.min_chunks_per_cta),
                                         **kw) for g, e, f in nets]
        self.streams = [torch.cuda.Stream() for _ in self.engines]
        self._next = 0
        self.scale = netG.scale

    def __len__(self):
        return len(self.engines)

    @property
    def launches_per_step(self):
        return self.engines[0].launches_per_step

    def warm(self, frames, B=1):
        """Capture every pipeline's CUDA graphs for this window shape (sequentially, on the current stream)."""
        for eng in self.engines:
            eng.adapt_and_infer_nhwc(frames, B)
        torch.cuda.current_stream().synchronize()

    def submit(self, fn, pipeline=None):
        """Run ``fn(engine)`` on the next pipeline's stream (round robin) after everything already enqueued on the
        caller's stream; returns (pipeline index, result of fn).  Nothing is synchronised: results are ordered on that
        pipeline's stream -- consume them there (``fn`` may include the device-to-host copy) or call ``join()``."""
        i = self._next if pipeline is None else pipeline
        self._next = (i + 1) % len(self.engines)
        s = self.streams[i]
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            out = fn(self.engines[i])
        return i, out

    def join(self):
        """Make the caller's stream wait for every pipeline."""
        cur = torch.cuda.current_stream()
        for s in self.streams:
            cur.wait_stream(s)

    def adapt_and_infer_many(self, windows):
        """windows: iterable of [N, H, W, 3] device tensors (LR windows, channels-last) -> list of HR [1, sH, sW, 3]."""
        outs = []
        for w in windows:
            w.record_stream(self.streams[self._next])           # allocator: w is consumed on the pipeline's stream
            _, hr = self.submit(lambda eng, w=w: eng.adapt_and_infer_nhwc(w).clone())
            outs.append(hr)
        self.join()
        for hr in outs:
            hr.record_stream(torch.cuda.current_stream())
        return outs

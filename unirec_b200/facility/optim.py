"""FusedOptimizer: the optimizer step of the hot path on the CUDA kernels.

Replaces `optim.Adam(model.parameters())` + `clip_grad_norm_` of the reference trainer
(unirec/facility/trainer.py:134-152, 347-349):
  * all dense (encoder / bias) parameters live in ONE flat buffer -> one `ur_dense_opt_f32` launch;
  * embedding tables are updated row-sparsely from the engine's row lists (`ur_rowlist_link` + `ur_rowlist_apply_f32`):
    only rows present in the batch are read and written, with Adam moments kept per row ("lazy" Adam = the dense
    Adam arithmetic restricted to touched rows; documented deviation H1, exact on step 1 and whenever every row
    with non-zero moments is touched);  `table_update: dense` switches tables to the exact dense update;
  * global-norm clipping uses a squared-norm pass over the same row lists, no dense gradient is ever built;
  * a NaN loss skips the whole update on the device (no host sync): trainer.py:344-352.
"""
import torch

from unirec_b200 import ops

_KERNEL_MODES = {'adam': 'adam', 'adamw': 'adamw', 'sgd': 'sgd'}


class FusedOptimizer(torch.optim.Optimizer):
    def __init__(self, model, opt_type='adam', lr=1e-3, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8):
        if opt_type == 'sparse_adam':
            opt_type = 'adam'
        if opt_type not in _KERNEL_MODES:
            raise ValueError("FusedOptimizer implements optimizer in ('adam','adamw','sgd','sparse_adam'); got %r. "
                             "Use table_update=dense with a torch optimizer for the others." % (opt_type,))
        params = [p for p in model.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, weight_decay=weight_decay, betas=betas, eps=eps))
        self.model = model
        self.mode = _KERNEL_MODES[opt_type]
        self.max_grad_norm = None
        self._st = None
        self._early = None          # (RowGrad, event) of the early table update issued for the current step

    # ---- state ----------------------------------------------------------------------------------
    def _state(self):
        eng = self.model._engine
        eng.ensure_ready()
        if self._st is None or self._st['flat'] is not eng.flat:
            dev = eng.device
            st = dict(flat=eng.flat, step=torch.zeros(1, dtype=torch.int32, device=dev),
                      sqnorm=torch.zeros(1, dtype=torch.float32, device=dev),
                      clip=torch.ones(1, dtype=torch.float32, device=dev), tables={})
            if self.mode != 'sgd':
                st['m'] = torch.zeros_like(eng.flat.data)
                st['v'] = torch.zeros_like(eng.flat.data)
            self._st = st
        return self._st

    def _table_state(self, st, p):
        ts = st['tables'].get(id(p))
        if ts is None:
            ts = {}
            if self.mode != 'sgd':
                ts['m'] = torch.zeros_like(p.data)
                ts['v'] = torch.zeros_like(p.data)
            st['tables'][id(p)] = ts
        return ts

    def zero_grad(self, set_to_none=True):
        self.model._engine.zero_dense_grads()
        for p in self.model._engine.table_params():
            p.grad = None

    # ---- early table update (overlaps the encoder backward) -------------------------------------
    def _hyper(self, st, scale=None):
        g = self.param_groups[0]
        (b1, b2) = g['betas']
        return dict(lr=g['lr'], beta1=b1, beta2=b2, eps=g['eps'], weight_decay=g['weight_decay'], step_dev=st['step'],
                    grad_scale_dev=scale, skip_flag=self.model._engine.nan_flag)

    @torch.no_grad()
    def early_apply(self, eng, rg, sources):
        """Called by the engine right after the loss kernel when the step's row lists were linked early: rows touched only by the
        targets/negatives (uniq[n_hist:]) have their complete gradient already (dL/ds x user embedding), so they are updated now,
        on the side stream, while the main stream runs the encoder backward.  step() later updates uniq[:n_hist]."""
        st = self._state()
        ts = self._table_state(st, rg.param)
        main = torch.cuda.current_stream()
        ev = eng._ev_main
        ev.record(main)                                # loss kernel + NaN flag are complete
        eng._side.wait_event(ev)
        with torch.cuda.stream(eng._side):
            ops.step_advance(st['step'], eng.nan_flag)
            ops.rowlist_apply(rg.param.data, ts.get('m'), ts.get('v'), rg.head, rg.next, rg.uniq, rg.n_uniq, sources[0][4],
                              sources, self.mode, u_begin=rg.n_hist, small_ctas=True, **self._hyper(st))
            done = torch.cuda.Event()
            done.record(eng._side)
        self._early = (rg, done)

    def _trainable_ranges(self, eng):
        """[(begin, end)] element ranges of the flat buffer covering the parameters with requires_grad (usually one range)."""
        named = dict(self.model.named_parameters())
        key = tuple(bool(named[n].requires_grad) for n in eng.flat.names)
        cached = getattr(self, '_ranges', None)
        if cached is not None and cached[0] == key:
            return cached[1]
        ranges = []
        for n, on in zip(eng.flat.names, key):
            off, cnt, _ = eng.flat.offsets[n]
            end = off + (cnt + 3) // 4 * 4
            if not on:
                continue
            if ranges and ranges[-1][1] == off:
                ranges[-1][1] = end
            else:
                ranges.append([off, end])
        ranges = [(a, min(b, eng.flat.size)) for a, b in ranges]
        self._ranges = (key, ranges)
        return ranges

    # ---- step -----------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        eng = self.model._engine
        st = self._state()
        g = self.param_groups[0]
        lr, wd, (b1, b2), eps = g['lr'], g['weight_decay'], g['betas'], g['eps']
        skip = eng.nan_flag
        dense_tables = self.model.table_update == 'dense'
        for rg in eng.rowgrads():
            if rg.specs and not rg.param.requires_grad and not rg.linked:      # frozen table: drop its gradient entries
                rg.reset()
        rowgrads = [rg for rg in eng.rowgrads() if rg.specs]
        early_rg = None
        for rg in rowgrads:
            if rg.early:                    # lists were linked on the side stream during the forward pass
                torch.cuda.current_stream().wait_event(eng._ev_link)
        if self._early is not None:
            early_rg, done = self._early
            self._early = None
            torch.cuda.current_stream().wait_event(done)
            if self.max_grad_norm is not None or dense_tables:
                raise RuntimeError('early table update is incompatible with gradient clipping / dense table updates '
                                   '(the Trainer disables it in those configurations)')
        dense_grads = {}
        if dense_tables:
            for rg in rowgrads:
                dense_grads[id(rg.param)] = rg.to_dense()
        else:
            for rg in rowgrads:
                rg.link()
        if early_rg is None:
            ops.step_advance(st['step'], skip)      # (already advanced by early_apply otherwise)

        scale = None
        if self.max_grad_norm is not None:
            st['sqnorm'].zero_()
            ops.sqnorm_accum(eng.flat.grad, st['sqnorm'])
            for rg in rowgrads:
                if dense_tables:
                    ops.sqnorm_accum(dense_grads[id(rg.param)].view(-1), st['sqnorm'])
                else:
                    ops.rowlist_apply(rg.param.data, None, None, rg.head, rg.next, rg.uniq, rg.n_uniq, rg.n_entries(),
                                      rg.sources(), 'sqnorm', sqnorm_out=st['sqnorm'])
            ops.clip_coef(st['sqnorm'], self.max_grad_norm, st['clip'])
            scale = st['clip']

        hyper = dict(lr=lr, beta1=b1, beta2=b2, eps=eps, weight_decay=wd, step_dev=st['step'], grad_scale_dev=scale,
                     skip_flag=skip)
        # frozen parameters (Trainer.load_model with `freeze`, unirec/facility/trainer.py:383-386: requires_grad = False) are left
        # alone: the flat buffer is updated in maximal runs of trainable parameters, frozen tables are skipped
        for o0, o1 in self._trainable_ranges(eng):
            ops.dense_opt(eng.flat.data[o0:o1], eng.flat.grad[o0:o1], st['m'][o0:o1] if 'm' in st else None,
                          st['v'][o0:o1] if 'v' in st else None, self.mode, **hyper)
        for p in eng.table_params():
            if not p.requires_grad and not eng.rowgrad(p).specs:
                continue
            ts = self._table_state(st, p)
            if dense_tables:
                # reference semantics: every row moves every step (zero gradient where untouched)
                gd = dense_grads.get(id(p))
                if gd is None:
                    gd = torch.zeros_like(p.data)
                ops.dense_opt(p.data.view(-1), gd.view(-1), ts.get('m').view(-1) if 'm' in ts else None,
                              ts.get('v').view(-1) if 'v' in ts else None, self.mode, **hyper)
            else:
                rg = eng.rowgrad(p)
                if rg.specs and rg is early_rg:
                    # remaining rows: the ones the history touches (their lists may also hold target entries)
                    ops.rowlist_apply(p.data, ts.get('m'), ts.get('v'), rg.head, rg.next, rg.uniq, rg.n_uniq,
                                      rg.specs[1][0].numel(), rg.sources(), self.mode, u_end=rg.n_hist, **hyper)
                elif rg.specs:
                    ops.rowlist_apply(p.data, ts.get('m'), ts.get('v'), rg.head, rg.next, rg.uniq, rg.n_uniq, rg.n_entries(),
                                      rg.sources(), self.mode, **hyper)
        for rg in rowgrads:
            rg.reset()
        return None

    # ---- checkpointing --------------------------------------------------------------------------
    def state_dict(self):
        st = self._st
        out = {'param_groups': [{k: v for k, v in g.items() if k != 'params'} for g in self.param_groups], 'mode': self.mode}
        if st is not None:
            out['step'] = st['step'].clone()
            for k in ('m', 'v'):
                if k in st:
                    out['flat_' + k] = st[k].clone()
            names = {id(p): n for n, p in self.model.named_parameters()}
            out['tables'] = {names[pid]: {k: v.clone() for k, v in ts.items()} for pid, ts in st['tables'].items()
                             if pid in names}
        return out

    def load_state_dict(self, sd):
        for g, s in zip(self.param_groups, sd.get('param_groups', [])):
            g.update(s)
        if 'step' not in sd:
            return
        st = self._state()
        st['step'].copy_(sd['step'])
        for k in ('m', 'v'):
            if 'flat_' + k in sd and k in st:
                st[k].copy_(sd['flat_' + k])
        params = dict(self.model.named_parameters())
        for n, ts in sd.get('tables', {}).items():
            if n in params:
                cur = self._table_state(st, params[n])
                for k, v in ts.items():
                    cur[k].copy_(v)

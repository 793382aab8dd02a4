"""torch.optim.Adam-compatible optimizer whose step is the fused clip+Adam kernel.

The reference builds `torch.optim.Adam(bilinear.parameters(), lr)` (model/bilinear.py:60), writes
`param_groups[i]['lr']` (train_bilinear.py:69-70), calls zero_grad()/step() and round-trips
state_dict()/load_state_dict() through the checkpoint (train_bilinear.py:92-104, model/bilinear.py:82).
FusedAdam subclasses torch.optim.Adam so all of that keeps working; exp_avg / exp_avg_sq are views
into the engine's flat buffers and `step` is materialised in the layout torch expects.
"""
import torch

from .engine import get_engine


class FusedAdam(torch.optim.Adam):
    def __init__(self, module, lr=1.0e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(module.parameters(), lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False)
        self._module = module
        self._t = 0          # Adam's per-parameter step count (identical for all 22 tensors)
        self._adopted_for = None   # address of the flat moment buffer self.state was last checked against

    # ------------------------------------------------------------------ helpers
    def _hyper(self):
        g = self.param_groups[0]
        return g["lr"], g["betas"], g["eps"]

    def _adopt_state(self):
        """Point self.state[p] at the flat moment buffers (after the first step / after load)."""
        eng = get_engine(self._module)
        eng.ensure_flat()
        m, v = eng.state_views()
        for p, mi, vi in zip(self._module.parameters(), m, v):
            self.state[p] = {"step": torch.tensor(float(self._t)), "exp_avg": mi, "exp_avg_sq": vi}

    def _sync_foreign_state(self):
        """Moments that are NOT views of the engine's flat buffers -- a checkpoint loaded while the module
        was still on the CPU (load(device=cpu, parameter_dir=...) followed by .to('cuda')), or buffers
        re-created after a device change -- are copied in before the kernel reads them.  Without this the
        kernel would run Adam on zeroed moments with step count > 0 and the next state_dict() would
        overwrite what was loaded."""
        if self._t <= 0 or not self.state:
            return
        eng = get_engine(self._module)
        eng.ensure_flat()
        if self._adopted_for == eng.flat_m.data_ptr():     # hot path: one integer comparison per step
            return
        m, v = eng.state_views()
        foreign = False
        with torch.no_grad():
            for p, mi, vi in zip(self._module.parameters(), m, v):
                st = self.state.get(p)
                if not st or "exp_avg" not in st:
                    continue
                if st["exp_avg"].data_ptr() != mi.data_ptr() or st["exp_avg_sq"].data_ptr() != vi.data_ptr():
                    mi.copy_(st["exp_avg"].to(mi.device).reshape(mi.shape))
                    vi.copy_(st["exp_avg_sq"].to(vi.device).reshape(vi.shape))
                    foreign = True
        if foreign:
            self._adopt_state()
        self._adopted_for = eng.flat_m.data_ptr()

    # ------------------------------------------------------------------ torch.optim API
    @torch.no_grad()
    def step(self, closure=None):
        """optimizer.step() of train_bilinear.py:83 on the gradients currently in param.grad.  Gradient
        clipping, if any, was already applied by the caller (nn.utils.clip_grad_norm_)."""
        if closure is not None:
            raise RuntimeError("closures are not supported")
        eng = get_engine(self._module)
        eng.ensure_flat()
        self._sync_foreign_state()
        views = eng.grad_views()
        for p, v in zip(self._module.parameters(), views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
        lr, betas, eps = self._hyper()
        self._t += 1
        eng.adam_step(eng.adam_args(lr, self._t, 0.0, 1.0, betas, eps))
        self._adopt_state()

    def fused_step(self, module, x, y, max_norm=1.0, update=True, grad_scale=1.0):
        assert module is self._module
        eng = get_engine(module)
        self._sync_foreign_state()
        lr, betas, eps = self._hyper()
        if update:
            self._t += 1
        out = eng.train_step(x, y, eng.adam_args(lr, max(self._t, 1), max_norm, grad_scale, betas, eps), update=update)
        if update and not self.state:
            self._adopt_state()
        return out

    def fused_step_dp(self, module, x, y, max_norm=1.0, world=1):
        assert module is self._module
        eng = get_engine(module)
        self._sync_foreign_state()
        lr, betas, eps = self._hyper()
        self._t += 1
        out = eng.train_step_dp(x, y, eng.adam_args(lr, self._t, max_norm, 1.0 / world, betas, eps))
        if not self.state:
            self._adopt_state()
        return out

    def fused_step_host(self, module, x_host, y_host, max_norm=1.0, update=True, grad_scale=1.0):
        assert module is self._module
        eng = get_engine(module)
        self._sync_foreign_state()
        lr, betas, eps = self._hyper()
        if update:
            self._t += 1
        out = eng.train_step_host(x_host, y_host, eng.adam_args(lr, max(self._t, 1), max_norm, grad_scale, betas, eps),
                                  update=update)
        if update and not self.state:
            self._adopt_state()
        return out

    def finish_step(self, max_norm=1.0, grad_scale=1.0):
        """Data parallel: clip + Adam on the (already all-reduced) flat gradient buffer."""
        eng = get_engine(self._module)
        self._sync_foreign_state()
        lr, betas, eps = self._hyper()
        self._t += 1
        out = eng.adam_step(eng.adam_args(lr, self._t, max_norm, grad_scale, betas, eps))
        if not self.state:
            self._adopt_state()
        return out

    def state_dict(self):
        # on the device the moments live in the engine's flat buffers; a module still on the CPU (checkpoint
        # loaded, not yet moved) holds them as ordinary torch.optim.Adam state -- nothing to adopt there
        if self._t > 0 and next(self._module.parameters()).is_cuda:
            self._sync_foreign_state()
            self._adopt_state()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._adopted_for = None
        params = list(self._module.parameters())
        steps = [float(self.state[p]["step"]) for p in params if p in self.state and "step" in self.state[p]]
        self._t = int(steps[0]) if steps else 0
        if self._t > 0 and params[0].is_cuda:
            eng = get_engine(self._module)
            eng.ensure_flat()
            m, v = eng.state_views()
            with torch.no_grad():
                for p, mi, vi in zip(params, m, v):
                    st = self.state[p]
                    mi.copy_(st["exp_avg"])
                    vi.copy_(st["exp_avg_sq"])
            self._adopt_state()

"""ReparamModule: run a wrapped module with all of its parameters taken from one flat vector.

Same public surface as /root/reference/reparam_module.py:9-177 (``flat_param``, ``param_numel``,
``forward(*inputs, flat_param=, buffers=)``, ``embed(...)``): parameters are flattened in
``named_modules() x named_parameters(recurse=False)`` order (reference :30-51), removed from the
wrapped module, and re-installed as views of whichever flat tensor a call supplies, so autograd
reaches the flat vector (the MTT student, distill_s2d_ms.py:197-266).
"""
from contextlib import contextmanager

import torch
import torch.nn as nn


class ReparamModule(nn.Module):
    def __init__(self, module):
        super().__init__()
        self.module = module
        slots, shared, seen, tensors = [], [], {}, []
        for mod_name, mod in self.named_modules():
            for pname, p in mod.named_parameters(recurse=False):
                if p is None:
                    continue
                if p in seen:
                    shared.append((mod_name, pname) + seen[p])
                    continue
                seen[p] = (mod_name, pname)
                slots.append((mod_name, pname))
                tensors.append(p.detach())
        assert len({t.dtype for t in tensors}) <= 1, 'expects all parameters in module to have same dtype'
        self._param_infos = tuple(slots)
        self._shared_param_infos = tuple(shared)
        self._param_numels = tuple(t.numel() for t in tensors)
        self._param_shapes = tuple(t.size() for t in tensors)
        self.register_parameter('flat_param', nn.Parameter(torch.cat([t.reshape(-1) for t in tensors], 0)))
        self.param_numel = self.flat_param.numel()
        for mod_name, pname in self._param_infos:
            delattr(self._resolve(mod_name), pname)
        for mod_name, pname, _, _ in self._shared_param_infos:
            delattr(self._resolve(mod_name), pname)
        self._install(self.flat_param)
        self._buffer_infos = tuple((mn, bn, b) for mn, m in self.named_modules()
                                   for bn, b in m.named_buffers(recurse=False) if b is not None)

    # reference name kept for drop-in use
    def _get_module_from_name(self, mn):
        return self._resolve(mn)

    def _resolve(self, dotted):
        m = self
        if dotted:
            for part in dotted.split('.'):
                m = getattr(m, part)
        return m

    def _install(self, flat):
        """Set every parameter slot to a view of ``flat`` (plain attributes, not Parameters)."""
        views = (t.view(s) for t, s in zip(flat.split(self._param_numels), self._param_shapes))
        for (mod_name, pname), v in zip(self._param_infos, views):
            setattr(self._resolve(mod_name), pname, v)
        for mod_name, pname, src_mod, src_name in self._shared_param_infos:
            setattr(self._resolve(mod_name), pname, getattr(self._resolve(src_mod), src_name))

    _unflatten_param = _install

    def clear_views(self):
        for mod_name, pname in self._param_infos:
            setattr(self._resolve(mod_name), pname, None)

    @contextmanager
    def unflattened_param(self, flat_param):
        saved = [getattr(self._resolve(mn), pn) for mn, pn in self._param_infos]
        self._install(flat_param)
        try:
            yield
        finally:
            for (mn, pn), v in zip(self._param_infos, saved):
                setattr(self._resolve(mn), pn, v)
            for mn, pn, smn, spn in self._shared_param_infos:
                setattr(self._resolve(mn), pn, getattr(self._resolve(smn), spn))

    @contextmanager
    def replaced_buffers(self, buffers):
        for (mn, bn, _), nb in zip(self._buffer_infos, buffers):
            setattr(self._resolve(mn), bn, nb)
        try:
            yield
        finally:
            for mn, bn, old in self._buffer_infos:
                setattr(self._resolve(mn), bn, old)

    def _call(self, fn_name, inputs, kwinputs, flat_param, buffers):
        # like the reference (:149), a (1, P) row scattered by DataParallel is squeezed to (P,)
        flat_param = self.flat_param if flat_param is None else torch.squeeze(flat_param)
        fn = self.module if fn_name is None else getattr(self.module, fn_name)
        with self.unflattened_param(flat_param):
            if buffers is None:
                return fn(*inputs, **kwinputs)
            with self.replaced_buffers(tuple(buffers)):
                return fn(*inputs, **kwinputs)

    def forward(self, *inputs, flat_param=None, buffers=None, **kwinputs):
        return self._call(None, inputs, kwinputs, flat_param, buffers)

    def embed(self, *inputs, flat_param=None, buffers=None, **kwinputs):
        return self._call('embed', inputs, kwinputs, flat_param, buffers)

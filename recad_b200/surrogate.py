"""WMF surrogate trainer of the AIA / Leg-UP attack loop on the CUDA path: drop-in for
recad/model/attacker/aia.py:393-489 `WMFTrainer` (same constructor arguments, same `fit_adv(data_tensor, epoch_num,
unroll_steps)` returning the surrogate's predictions with a graph back to `data_tensor`).

The reference retrains this model from scratch once per attack step (aia.py:88-114; 50 x per step in
aushplus.py:175-178): `epoch_num - unroll_steps` plain Adam epochs, then `unroll_steps` epochs inside
`higher.innerloop_ctx` so that the attacker's loss can be differentiated through the training of the surrogate.  Here
both phases are ONE persistent cluster kernel each (csrc/wmf.cu) and the reverse pass through the unrolled Adam steps is
a third; `higher` is not needed.  Host work per call: the initial factors (torch CPU generator, as the reference) and
the epoch shuffles (np.random.shuffle on the global stream, as the reference, aia.py:447 / 468).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, ops


class _UnrolledWMF(torch.autograd.Function):
    """(data) -> (P_final, Q_final) of WMFTrainer.fit_adv; backward = the reverse pass through the unrolled epochs."""

    @staticmethod
    def forward(ctx, data, trainer, epoch_num, unroll_steps):
        st, keep, orders, snap, n_plain = trainer._run(data.detach(), epoch_num, unroll_steps, trainer._init, trainer._orders)
        ctx.trainer, ctx.st, ctx.keep, ctx.orders, ctx.snap = trainer, st, keep, orders, snap
        ctx.n_plain, ctx.unroll_steps = n_plain, unroll_steps
        ctx.save_for_backward(data)
        P, Q = keep[0], keep[1]
        return P.clone(), Q.clone()

    @staticmethod
    def backward(ctx, Pbar, Qbar):
        (data,) = ctx.saved_tensors
        if ctx.unroll_steps == 0 or ctx.snap is None:
            return torch.zeros_like(data), None, None, None
        tr, st = ctx.trainer, ctx.st
        dev = data.device
        spe = (st.n_rows + st.batch - 1) // st.batch
        with torch.cuda.device(dev):
            Pb, Qb = Pbar.contiguous().clone(), Qbar.contiguous().clone()
            scratch = torch.empty(2 * (st.n_rows + st.n_items) * st.dim + 1024, dtype=torch.float32, device=dev)
            d_data = torch.zeros_like(data)
            o = ctx.orders[ctx.n_plain:].contiguous()
            _lib.check(_lib.lib().recad_wmf_backward(C.byref(st), data.data_ptr(), o.data_ptr(), ctx.unroll_steps, ctx.n_plain * spe,
                                                     ctx.snap.data_ptr(), Pb.data_ptr(), Qb.data_ptr(), scratch.data_ptr(),
                                                     d_data.data_ptr(), ops._stream(dev)), "recad_wmf_backward")
        return d_data, None, None, None


class WMFTrainer:
    def __init__(self, n_users, n_items, device, hidden_dim, lr, weight_decay, batch_size, weight_pos, weight_neg, verbose=False):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ops.RecadError("recad_b200.surrogate.WMFTrainer computes on a CUDA device only; there is no CPU fallback")
        self.n_users, self.n_items, self.hidden_dim = int(n_users), int(n_items), int(hidden_dim)
        self.lr, self.weight_decay, self.batch_size = float(lr), float(weight_decay), int(batch_size)
        self.weight_pos, self.weight_neg, self.verbose = float(weight_pos), float(weight_neg), verbose
        self.P = self.Q = None
        self._init = self._orders = None

    def _initialize(self):
        """WeightedMF.__init__ (aia.py:230-236): Q then P from the global torch CPU generator, N(0, 0.1)."""
        Q = torch.zeros([self.n_items, self.hidden_dim]).normal_(mean=0, std=0.1)
        P = torch.zeros([self.n_users, self.hidden_dim]).normal_(mean=0, std=0.1)
        return P, Q

    def _run(self, data, epoch_num, unroll_steps, init=None, orders=None):
        dev = self.device
        n_rows, n_items = int(data.shape[0]), int(data.shape[1])
        if n_rows != self.n_users or n_items != self.n_items:
            raise ValueError(f"data_tensor is {tuple(data.shape)}, the trainer was built for {(self.n_users, self.n_items)}")
        P0, Q0 = init if init is not None else self._initialize()
        if orders is None:                                   # ONE idx_list shuffled in place every epoch (aia.py:443-447, 468)
            idx = np.arange(n_rows)
            orders = np.empty((epoch_num, n_rows), dtype=np.int32)
            for e in range(epoch_num):
                np.random.shuffle(idx)
                orders[e] = idx
        with torch.cuda.device(dev):
            P, Q = P0.to(dev, torch.float32).contiguous().clone(), Q0.to(dev, torch.float32).contiguous().clone()
            mP, vP, mQ, vQ = (torch.zeros_like(t) for t in (P, P, Q, Q))
            o = torch.as_tensor(np.ascontiguousarray(orders, dtype=np.int32)).to(dev)
            data = data.to(dev, torch.float32).contiguous()
            st = _lib.WMF()
            st.n_rows, st.n_items, st.dim, st.batch = n_rows, n_items, self.hidden_dim, self.batch_size
            st.lr, st.beta1, st.beta2, st.eps = self.lr, 0.9, 0.999, 1e-8
            st.weight_decay, st.weight_pos, st.weight_neg = self.weight_decay, self.weight_pos, self.weight_neg
            st.P, st.Q, st.mP, st.vP, st.mQ, st.vQ = (t.data_ptr() for t in (P, Q, mP, vP, mQ, vQ))
            L = _lib.lib()
            n_plain = epoch_num - unroll_steps
            spe = (n_rows + self.batch_size - 1) // self.batch_size
            if n_plain > 0:
                _lib.check(L.recad_wmf_fit(C.byref(st), data.data_ptr(), o.data_ptr(), n_plain, 0, 0, None, ops._stream(dev)), "recad_wmf_fit")
            snap = None
            if unroll_steps > 0:
                snap = torch.empty(L.recad_wmf_snapshot_floats(C.byref(st), unroll_steps), dtype=torch.float32, device=dev)
                _lib.check(L.recad_wmf_fit(C.byref(st), data.data_ptr(), o[n_plain:].contiguous().data_ptr(), unroll_steps, n_plain * spe, 1,
                                           snap.data_ptr(), ops._stream(dev)), "recad_wmf_fit")
        self.P, self.Q = P, Q
        return st, (P, Q, mP, vP, mQ, vQ), o, snap, n_plain

    def fit_adv(self, data_tensor, epoch_num, unroll_steps, init=None, orders=None):
        """aia.py:431-489.  Returns P Q^T [n_rows, n_items]; gradients flow to `data_tensor` through the unrolled epochs.
        init = (P0, Q0) / orders = int [epoch_num, n_rows] replace the draws from the torch / numpy generators (tests)."""
        if not data_tensor.requires_grad:
            raise ValueError("To compute adversarial gradients, data_tensor should have requires_grad=True.")
        self._init, self._orders = init, orders
        try:
            P, Q = _UnrolledWMF.apply(data_tensor.to(self.device), self, int(epoch_num), int(unroll_steps))
        finally:
            self._init = self._orders = None
        return torch.mm(P, Q.t())

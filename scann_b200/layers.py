"""Layer-level mirror of the reference's Keras layers (scann/layers/attention.py,
scann/layers/custom_layers.py, scann/layers/losses.py) on top of the sm_100a kernels.

Same constructor arguments, ``get_config`` keys, weight order and call signatures as the
reference classes, with padded ``[B,M,N,...]`` tensors in and out, so parity tests read like
tests of the reference layers.  Tensors may be numpy arrays, torch CUDA tensors or anything
exporting ``__dlpack__`` (what ``tf.experimental.dlpack.to_dlpack`` hands over); outputs are
torch CUDA tensors (``torch.utils.dlpack.to_dlpack`` gives the capsule TF would consume).

Only the configuration the reference's ``create_model`` actually builds is accelerated
(``v_proj=False, kq_proj=True``; scann_model.py:395-403, :432-434); anything else raises.
Deviation, documented in DESIGN.md: masked neighbour slots are never computed, so the returned
``neighbor_geometry`` holds zeros there (the reference holds values no model output depends on).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import _abi
from ._abi import check, lib, ptr_array
from .dlpack import import_tensor

D = 128
TILE = 128
PLAN_GSZ = 128


def _dev() -> torch.device:
    _abi.require_gpu()
    return torch.device("cuda", torch.cuda.current_device())


def _t(x, dtype=torch.float32) -> torch.Tensor:
    x = import_tensor(x)                         # DLPack capsule / __dlpack__ exporter -> torch view (zero-copy)
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    t = t.to(_dev())
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _p(t: Optional[torch.Tensor], off: int = 0) -> int:
    return 0 if t is None else t.data_ptr() + off * t.element_size()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# --------------------------------------------------------------------------- plain functions
def gather_shape(x):
    """gather_shape (custom_layers.py:18-28): [B,M,N] -> [B,M,N,2] = (batch id, neighbour id).
    Pure index bookkeeping; the fused kernels take the neighbour ids directly."""
    x = _t(x, torch.int32)
    B, M, N = x.shape
    rb = torch.arange(B, dtype=torch.int32, device=x.device).view(B, 1, 1, 1).expand(B, M, N, 1)
    return torch.cat([rb, x.unsqueeze(-1)], -1)


def mrelu(x):
    """mrelu forward (custom_layers.py:6-15); its identity gradient lives in the head backward kernel."""
    return torch.clamp_min(_t(x), 0.0)


def root_mean_squared_error(y_true, y_pred):
    """losses.py:5-6 on host arrays (metric reporting only; the training loss is computed on device)."""
    y_true, y_pred = np.asarray(y_true, np.float64), np.asarray(y_pred, np.float64)
    return float(np.sqrt(np.mean((y_pred - y_true) ** 2)))


def r2_square(y_true, y_pred):
    """losses.py:13-16 (K.epsilon() = 1e-7)."""
    y_true, y_pred = np.asarray(y_true, np.float64), np.asarray(y_pred, np.float64)
    ss_res = np.sum((y_true - y_pred) ** 2)
    ss_tot = np.sum((y_true - np.mean(y_true)) ** 2)
    return float(1 - ss_res / (ss_tot + 1e-7))


class GaussianExpansion:
    """GaussianExpansion(centers, width=0.5) (custom_layers.py:31-75).  Inside the model the
    expansion is fused into the geometry-initialisation kernel; this standalone form exists for
    API parity and runs as a torch elementwise op."""

    def __init__(self, centers, width=0.5, **kwargs):
        self.centers = np.asarray(centers)
        self.width = float(np.diff(self.centers).mean()) if width is None else width ** 2

    def __call__(self, inputs, masks=None):
        d = _t(inputs)
        c = torch.from_numpy(self.centers.astype(np.float32)).to(d.device)
        return torch.exp(-((d.unsqueeze(-1) - c) ** 2) / self.width)

    call = __call__

    def get_config(self):
        return {"centers": self.centers}


# --------------------------------------------------------------------------- pair plan for a layer call
class _Plan:
    def __init__(self, neighbors: torch.Tensor, mask_u8: torch.Tensor):
        dev = neighbors.device
        B, M, N = neighbors.shape
        self.B, self.M, self.N, self.R = B, M, N, B * M
        P = B * M * N
        ngroups = (self.R + PLAN_GSZ - 1) // PLAN_GSZ
        cap = P // (129 - N) + ngroups + 1 if N <= 64 else 2 * (P // TILE) + ngroups + 2
        self.cap = cap
        rows = cap * TILE
        i32 = dict(dtype=torch.int32, device=dev)
        self.cnt = torch.empty(self.R, **i32)
        self.rowptr = torch.empty(self.R, **i32)
        self.tile_a0 = torch.empty(cap, **i32)
        self.tile_a1 = torch.empty(cap, **i32)
        self.ntiles = torch.zeros(1, **i32)
        self.pair_c = torch.empty(rows, **i32)
        self.pair_j = torch.empty(rows, **i32)
        self.pair_slot = torch.empty(rows, **i32)
        self.pair_d = torch.empty(rows, dtype=torch.float32, device=dev)
        self.pair_w = torch.empty(rows, dtype=torch.float32, device=dev)
        self.status = torch.zeros(1, **i32)
        scratch = torch.zeros(2 * ngroups + 8, **i32)      # the one-launch plan form relies on a zeroed buffer
        zeros = torch.zeros(P, dtype=torch.float32, device=dev)
        check(lib.scann_plan_build(_p(mask_u8), _p(neighbors), _p(zeros), _p(zeros), B, M, N, cap, TILE, TILE, _p(self.cnt),
                                   _p(self.rowptr), _p(self.tile_a0), _p(self.tile_a1), _p(self.ntiles),
                                   _p(self.pair_c), _p(self.pair_j), _p(self.pair_slot), _p(self.pair_d),
                                   _p(self.pair_w), 0, 0, 0, _p(scratch), scratch.numel(), _p(self.status), _stream()),
              "plan_build")
        s = int(self.status.item())
        if s:
            raise _abi.ScannAbiError(f"plan failed, device status {s}")
        self.valid = self.pair_c >= 0
        self.slot = self.pair_slot.long()[self.valid]


class _Weighted:
    weight_names: List[str] = []

    def set_weights(self, weights) -> None:
        weights = [np.asarray(w, np.float32) for w in weights]
        if len(weights) != len(self.weight_names):
            raise ValueError(f"expected {len(self.weight_names)} weight arrays ({self.weight_names})")
        self._w = {n: _t(w) for n, w in zip(self.weight_names, weights)}

    def get_weights(self):
        return [self._w[n].cpu().numpy() for n in self.weight_names]


class LocalAttention(_Weighted):
    """LocalAttention (attention.py:53-231)."""

    def __init__(self, dim=128, num_head=8, v_proj=True, scale=0.5, activation="swish", kq_proj=True, dropout=False,
                 g_update=False, **kwargs):
        self.dim, self.num_head, self.hdim = dim, num_head, dim // num_head
        self.scale, self.v_proj, self.kq_proj, self.dropout, self.g_update = scale, v_proj, kq_proj, dropout, g_update
        self.name = kwargs.get("name", "local_attention")
        if v_proj or not kq_proj or not g_update or dim != D or num_head != 8 or scale != 0.5 or activation != "swish":
            raise NotImplementedError("accelerated LocalAttention supports the configuration create_model builds: "
                                      "dim=128, num_head=8, v_proj=False, kq_proj=True, g_update=True, swish")
        self.weight_names = ["query/kernel", "query/bias", "key/kernel", "key/bias", "filter_geo/kernel",
                             "filter_geo/bias", "layer_norm/gamma", "layer_norm/beta", "layer_norm_g/gamma",
                             "layer_norm_g/beta"]
        self._w = {}

    def get_config(self):                                                    # attention.py:218-231
        return {"name": self.name, "dim": self.dim, "scale": self.scale, "num_head": self.num_head,
                "v_proj": self.v_proj, "kq_proj": self.kq_proj, "g_update": self.g_update, "dropout": self.dropout}

    def __call__(self, atom_query, atom_neighbor, neighbor_geometry, mask, neighbor_weight=None):
        """(attn[B,H,M,N], context[B,M,D], neighbor_geometry[B,M,N,D]) -- attention.py:118-216."""
        w = self._w
        x = _t(atom_query)
        idx = _t(atom_neighbor, torch.int32)
        B, M, N = idx.shape[:3]
        nbr = idx[..., 1].contiguous()
        m_u8 = (_t(mask, None) != 0).view(torch.uint8).contiguous()
        geom = _t(neighbor_geometry).reshape(B * M * N, D)
        plan = _Plan(nbr, m_u8)
        R, rows, dev = B * M, plan.cap * TILE, x.device
        g_in = torch.zeros(rows, D, dtype=torch.float32, device=dev)
        g_in[plan.valid] = geom[plan.slot]
        x2 = x.reshape(R, D).contiguous()
        proj = torch.empty(R, 3 * D, dtype=torch.float32, device=dev)
        fg = w["filter_geo/kernel"]
        check(lib.scann_dense_forward(ptr_array([_p(x2)]), D,
                                      ptr_array([_p(fg), _p(fg, 2 * D * D), _p(w["query/kernel"])]),
                                      ptr_array([_p(w["filter_geo/bias"]), 0, _p(w["query/bias"])]), 1, 3, R, _p(proj),
                                      3 * D, 0, 0, D, 0, 0, 0, 0, _stream()), "dense_forward")
        out = torch.empty(R, D, dtype=torch.float32, device=dev)
        g_out = torch.empty(rows, D, dtype=torch.float32, device=dev)
        attn_p = torch.zeros(rows, 8, dtype=torch.float32, device=dev)
        check(lib.scann_la_nopair_forward(_p(plan.cnt), _p(proj), R, _p(w["layer_norm/gamma"]),
                                          _p(w["layer_norm/beta"]), 0, _p(out), _stream()), "la_nopair")
        check(lib.scann_la_forward(_abi.require_gpu(), _p(plan.ntiles), _p(plan.tile_a0), _p(plan.tile_a1),
                                   _p(plan.cnt), _p(plan.rowptr), _p(plan.pair_c), _p(plan.pair_j), _p(x2), _p(proj),
                                   _p(g_in), _p(fg, D * D), _p(w["key/kernel"]), _p(w["key/bias"]),
                                   _p(w["layer_norm_g/gamma"]), _p(w["layer_norm_g/beta"]), _p(w["layer_norm/gamma"]),
                                   _p(w["layer_norm/beta"]), _p(g_out), 0, _p(out), _p(attn_p), _stream()),
              "la_forward")
        geom_out = torch.zeros(B * M * N, D, dtype=torch.float32, device=dev)
        geom_out[plan.slot] = g_out[plan.valid]
        # attn: valid slots from the kernel; masked slots are exactly 0 in fp32; rows without any valid
        # slot are the uniform softmax of N equal logits (attention.py:186-189)
        attn = torch.zeros(B * M * N, 8, dtype=torch.float32, device=dev)
        attn[plan.slot] = attn_p[plan.valid]
        attn = attn.reshape(B, M, N, 8)
        empty = (plan.cnt.reshape(B, M) == 0)
        attn[empty] = 1.0 / N
        return attn.permute(0, 3, 1, 2).contiguous(), out.reshape(B, M, D), geom_out.reshape(B, M, N, D)

    call = __call__


class ResidualNorm(_Weighted):
    """ResidualNorm (attention.py:19-50), inference form (Dropout inactive)."""

    def __init__(self, dim=128, dropout_rate=0.1, **kwargs):
        if dim != D:
            raise NotImplementedError("accelerated ResidualNorm supports dim=128")
        self.dim, self.dropout = dim, dropout_rate
        self.weight_names = ["dense/kernel", "dense/bias", "dense_1/kernel", "dense_1/bias", "layer_norm/gamma",
                             "layer_norm/beta"]
        self._w = {}

    def get_config(self):                                                    # attention.py:42-50
        return {"dim": self.dim, "dropout": self.dropout}

    def __call__(self, x):
        w = self._w
        x = _t(x)
        shape = x.shape
        x2 = x.reshape(-1, D).contiguous()
        R = x2.shape[0]
        h1 = torch.empty_like(x2)
        out = torch.empty_like(x2)
        check(lib.scann_dense_forward(ptr_array([_p(x2)]), D, ptr_array([_p(w["dense/kernel"])]),
                                      ptr_array([_p(w["dense/bias"])]), 1, 1, R, _p(h1), D, 1, 0, D, 0, 0, 0, 0,
                                      _stream()), "dense_forward")
        check(lib.scann_dense_forward(ptr_array([_p(h1)]), D, ptr_array([_p(w["dense_1/kernel"])]),
                                      ptr_array([_p(w["dense_1/bias"])]), 1, 1, R, _p(out), D, 3, _p(x2), D, 0, 0,
                                      _p(w["layer_norm/gamma"]), _p(w["layer_norm/beta"]), _stream()), "dense_forward")
        return out.reshape(shape)

    call = __call__


class GlobalAttention(_Weighted):
    """GlobalAttention (attention.py:234-331).  Returns (attn[B,M,1], context[B,D])."""

    def __init__(self, dim=128, v_proj=False, kq_proj=True, norm=True, **kwargs):
        if v_proj or not kq_proj or dim != D:
            raise NotImplementedError("accelerated GlobalAttention supports dim=128, v_proj=False, kq_proj=True")
        self.dim, self.norm, self.v_proj, self.kq_proj = dim, norm, v_proj, kq_proj
        self.name = "global_attention"                                       # ctor drops kwargs (attention.py:250)
        self.weight_names = ["query/kernel", "query/bias", "key/kernel", "key/bias"]
        self._w = {}

    def get_config(self):                                                    # attention.py:320-331
        return {"dim": self.dim, "norm": self.norm, "v_proj": self.v_proj, "kq_proj": self.kq_proj}

    def __call__(self, atom_query, mask):
        w = self._w
        x = _t(atom_query)
        B, M, _ = x.shape
        R = B * M
        x2 = x.reshape(R, D).contiguous()
        m_u8 = (_t(mask, None).reshape(R) != 0).view(torch.uint8).contiguous()
        dev = x.device
        qk = torch.empty(R, 2 * D, dtype=torch.float32, device=dev)
        check(lib.scann_dense_forward(ptr_array([_p(x2)]), D, ptr_array([_p(w["query/kernel"]), _p(w["key/kernel"])]),
                                      ptr_array([_p(w["query/bias"]), _p(w["key/bias"])]), 1, 2, R, _p(qk), 2 * D, 0, 0,
                                      D, 0, 0, 0, 0, _stream()), "dense_forward")
        ga = torch.empty(R, dtype=torch.float32, device=dev)
        y = torch.empty(B, dtype=torch.float32, device=dev)
        ctx = torch.empty(B, D, dtype=torch.float32, device=dev)
        zeros = torch.zeros(D * D + 2 * D + 1, dtype=torch.float32, device=dev)   # unused head weights
        check(lib.scann_ga_head_forward(_p(qk), _p(m_u8), B, M, int(self.norm), _p(zeros), _p(zeros, D * D),
                                        _p(zeros, D * D + D), _p(zeros, D * D + 2 * D), 0, _p(ga), _p(y), _p(ctx), 0,
                                        _stream()), "ga_head_forward")
        return ga.reshape(B, M, 1), ctx

    call = __call__

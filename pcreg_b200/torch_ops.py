"""Device-resident entry points: torch is only the plumbing (device memory, streams,
torch.distributed); the compute is libpcreg_b200.so through pcreg_icp_batch_dev."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .api import Model, icp_opts


def T_to_abi_t(T: torch.Tensor) -> torch.Tensor:
    """[H,4,4] math layout -> contiguous MATLAB column-major 16-double records."""
    return T.transpose(-1, -2).contiguous()


def T_from_abi_t(buf: torch.Tensor) -> torch.Tensor:
    return buf.view(-1, 4, 4).transpose(-1, -2).contiguous()


def src_to_abi_t(src: torch.Tensor) -> torch.Tensor:
    """[ns,3] -> column-major float64 (x[ns], y[ns], z[ns]) on the same device."""
    return src.to(torch.float64).t().contiguous()


class IcpDeviceBuffers:
    """Output tensors of one batched ICP call, allocated once and reused across steps."""

    def __init__(self, nhyp: int, ns: int, iters: int, device, want_idx=False, want_hist=False):
        f64, i32 = torch.float64, torch.int32
        self.T = torch.empty((nhyp, 16), dtype=f64, device=device)
        self.rmse = torch.empty(nhyp, dtype=f64, device=device)
        self.n_used = torch.empty(nhyp, dtype=i32, device=device)
        self.status = torch.empty(nhyp, dtype=i32, device=device)
        self.idx = torch.empty((nhyp, ns), dtype=i32, device=device) if want_idx else None
        self.hist = torch.empty((nhyp, iters + 1), dtype=f64, device=device) if want_hist else None
        self.best = torch.empty(1, dtype=torch.int64, device=device)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def icp_batch_device(model: Model, src_cm: torch.Tensor, w_src, T0_abi: torch.Tensor, opts, out: IcpDeviceBuffers,
                     stream=None):
    """Run pcreg_icp_batch_dev on tensors already resident on the library's device.

    src_cm: float64 [3, ns] (column-major ns x 3); T0_abi: float64 [H,16] column-major records."""
    if not (src_cm.is_cuda and T0_abi.is_cuda):
        raise L.PcregError("icp_batch_device needs CUDA tensors (there is no CPU fallback)")
    assert src_cm.dtype == torch.float64 and src_cm.is_contiguous() and src_cm.shape[0] == 3
    assert T0_abi.dtype == torch.float64 and T0_abi.is_contiguous()
    ns = src_cm.shape[1]
    nhyp = T0_abi.shape[0]
    st = stream if stream is not None else torch.cuda.current_stream()
    L.check(L.lib().pcreg_icp_batch_dev(model.handle, _p(src_cm), ns, _p(w_src), _p(T0_abi), nhyp, C.byref(opts),
                                        _p(out.T), _p(out.rmse), _p(out.n_used), _p(out.status), _p(out.idx),
                                        _p(out.hist), _p(out.best), C.c_void_p(st.cuda_stream)),
            "pcreg_icp_batch_dev")
    return out

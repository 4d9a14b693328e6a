"""CFG Euler/ODE sampler loop of ``sample_fitv2_ddp.py:273-314`` driven through the C ABI (both branches of
``using_cfg``: with guidance the fused CFG + Euler kernel, without it the plain Euler update).

One step = one ``fitv2_forward`` over 2n rows (the ``cat([z, z])`` of the script is done implicitly by
the patch-embed kernel) + one fused ``fitv2_cfg_euler`` kernel (CFG combine over all 16 channels and
the Euler update, fp32, bit-exact with the script's expression).  With ``use_cuda_graph`` the whole step
is captured once and replayed; the per-step timestep and step size live in device buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .model import FiT


def make_grid(n: int, n_patch_h: int, n_patch_w: int, device=None) -> torch.Tensor:
    """sample_fitv2_ddp.py:263-268 — (n, 2, N) int64, [:,0] = w index, [:,1] = h index."""
    gh = torch.arange(n_patch_h, dtype=torch.long)
    gw = torch.arange(n_patch_w, dtype=torch.long)
    g = torch.meshgrid(gw, gh, indexing="xy")
    grid = torch.cat([g[0].reshape(1, -1), g[1].reshape(1, -1)], dim=0).repeat(n, 1, 1)
    return grid.to(device) if device is not None else grid


class EulerCFGSampler:
    """Holds the static per-batch tensors (labels, grid, mask, sigma schedule) of one sampling job."""

    def __init__(self, model: FiT, y: torch.Tensor, grid: torch.Tensor, mask: torch.Tensor, num_steps: int,
                 cfg_scale: float, use_cuda_graph: bool = False, size: Optional[torch.Tensor] = None):
        self.model, self.num_steps, self.cfg_scale = model, num_steps, float(cfg_scale)
        self.using_cfg = cfg_scale > 1.0                                # sample_fitv2_ddp.py:242 (the script's default --cfg-scale is 1.0)
        dev = model.device
        n = y.shape[0]
        self.n = n
        if not model.use_sit:
            raise NotImplementedError("the Euler / CFG loop of sample_fitv2_ddp.py integrates a velocity model (use_sit=True)")
        if self.using_cfg and model.y_embedder.embedding_table.weight.shape[0] <= model.num_classes:
            raise ValueError("cfg_scale > 1 needs the null-class row of the label table (class_dropout_prob > 0, "
                             "sample_fitv2_ddp.py:277)")
        if model.online_rope and size is None:
            raise ValueError("online_rope=True: pass `size` (n, 1, 2) = (h, w) patches per sample (sample_fitv2_ddp.py:269-271)")
        # online_rope: the per-sample frequencies depend on `size` only, so they are bound ONCE here (duplicated for the CFG
        # rows like every other model input, sample_fitv2_ddp.py:278-282) instead of inside the step loop
        self.size2 = None
        if size is not None:
            self.size2 = (torch.cat([size, size], 0) if self.using_cfg else size).to(torch.int64)
        if self.using_cfg:
            self.y2 = torch.cat([y.to(dev, torch.int64), torch.full((n,), model.num_classes, dtype=torch.int64, device=dev)], 0)
            self.grid2 = torch.cat([grid, grid], 0).to(dev, torch.int64).contiguous()
            self.mask2 = torch.cat([mask, mask], 0).to(dev, torch.float32).contiguous()
        else:                                                           # :283-285, :300-301: the model sees the n rows as they are
            self.y2 = y.to(dev, torch.int64).contiguous()
            self.grid2 = grid.to(dev, torch.int64).contiguous()
            self.mask2 = mask.to(dev, torch.float32).contiguous()
        self.rows = self.y2.shape[0]
        sig = torch.linspace(0, 1, num_steps + 1)                       # CPU fp32, as the script (:287)
        self.sigmas = sig
        self.dsig = (sig[1:] - sig[:-1]).to(dev).contiguous()           # fp32 differences, per step
        self.t_table = sig[:-1, None].expand(num_steps, self.rows).to(dev).contiguous()
        self.step_scale = torch.stack([torch.ones(num_steps), sig[1:] - sig[:-1]], 1).to(dev).contiguous()   # (1, dsigma) per step
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._t_cur = torch.zeros(self.rows, dtype=torch.float32, device=dev)
        self._sc_cur = torch.zeros(2, dtype=torch.float32, device=dev)
        self._ds_cur = torch.zeros(1, dtype=torch.float32, device=dev)
        self._v2 = None
        self._z = None
        if model.online_rope:
            model._set_online_rope(self.size2, self.rows)

    def _step(self, z: torch.Tensor, t_rows: torch.Tensor, dsig_dev: torch.Tensor, scale_dev: Optional[torch.Tensor] = None):
        m = self.model
        l0 = m.kernel_launches()
        self._v2 = m._run(z, t_rows, self.y2, self.grid2, self.mask2, rows=self.rows, out=self._v2)
        self.launches_per_step = m.kernel_launches() - l0
        st = torch.cuda.current_stream(m.device).cuda_stream
        if not self.using_cfg:                                          # z = z + (sigma_next - sigma_current) * noise_pred  (:314), bit-exact
            if scale_dev is None:
                scale_dev = torch.stack([torch.ones_like(dsig_dev[0]), dsig_dev[0]])
            _lib.check(_lib.load().fitv2_scaled_add(C.c_void_p(z.data_ptr()), C.c_void_p(z.data_ptr()), C.c_void_p(self._v2.data_ptr()),
                                                    C.c_void_p(scale_dev.data_ptr()), z.numel(), C.c_void_p(st)), "fitv2_scaled_add")
            return
        _lib.check(_lib.load().fitv2_cfg_euler(C.c_void_p(z.data_ptr()), C.c_void_p(self._v2.data_ptr()), self.cfg_scale, 0.0,
                                               C.c_void_p(dsig_dev.data_ptr()), self.n, z.shape[1], z.shape[2], C.c_void_p(st)),
                   "fitv2_cfg_euler")

    @torch.no_grad()
    def sample(self, z: torch.Tensor, first_steps: Optional[int] = None) -> torch.Tensor:
        """z (n, N, C) fp32 on the model device; returns the latents after the trajectory (a new tensor)."""
        m = self.model
        z = z.to(m.device, torch.float32).contiguous().clone()
        steps = self.num_steps if first_steps is None else first_steps
        if m.online_rope:                                               # another caller may have bound other frequencies since
            m._set_online_rope(self.size2, self.rows)
        with torch.cuda.device(m.device):
            if not self.use_cuda_graph:
                for i in range(steps):
                    self._step(z, self.t_table[i], self.dsig[i:i + 1], self.step_scale[i])
                return z
            self.capture(z)
            self._z.copy_(z)
            for i in range(steps):
                self.replay_step(i)
            return self._z.clone()

    @torch.no_grad()
    def capture(self, z: torch.Tensor):
        """Capture one step (forward + update, 260-odd launches) into a CUDA graph operating on an internal latent buffer
        ``self._z`` shaped like ``z``; idempotent."""
        if self._graph is not None:
            return
        m = self.model
        with torch.cuda.device(m.device):
            self._z = torch.empty_like(z, dtype=torch.float32, device=m.device)
            self._z.copy_(z)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                               # warm-up outside capture (workspace, maps)
                self._t_cur.copy_(self.t_table[0]); self._ds_cur.zero_(); self._sc_cur.copy_(torch.tensor([1.0, 0.0]))
                self._step(self._z, self._t_cur, self._ds_cur, self._sc_cur)
            torch.cuda.current_stream().wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._step(self._z, self._t_cur, self._ds_cur, self._sc_cur)

    def replay_step(self, i: int):
        """One step of the captured graph (after ``sample`` has captured it): copies the step's scalars into the graph's
        device buffers and replays."""
        self._t_cur.copy_(self.t_table[i])
        self._ds_cur.copy_(self.dsig[i:i + 1])
        self._sc_cur.copy_(self.step_scale[i])
        self._graph.replay()


@torch.no_grad()
def euler_cfg_sample(model: FiT, z, y, grid, mask, size=None, num_steps: int = 250, cfg_scale: float = 1.5,
                     use_cuda_graph: bool = False, first_steps: Optional[int] = None) -> torch.Tensor:
    """Functional form of the script's loop: z (n,N,C), y (n,), grid (n,2,N), mask (n,N)."""
    return EulerCFGSampler(model, y, grid, mask, num_steps, cfg_scale, use_cuda_graph, size=size).sample(z, first_steps)


def pack_images_uint8(samples: torch.Tensor) -> torch.Tensor:
    """sample_fitv2_ddp.py:321-323 in one kernel: decoded images (B, C, H, W) fp32 on the GPU ->
    ``clamp(127.5 * samples.clamp(-1, 1) + 128, 0, 255).permute(0, 2, 3, 1).to(uint8)`` (B, H, W, C), bit-exact."""
    if not samples.is_cuda:
        raise _lib.FitV2Error("pack_images_uint8 needs a CUDA tensor: fitv2_b200 has no CPU path")
    x = samples.to(torch.float32).contiguous()
    B, Cc, H, W = x.shape
    out = torch.empty(B, H, W, Cc, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.load().fitv2_pack_uint8(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), B, Cc, H, W, C.c_void_p(st)),
                   "fitv2_pack_uint8")
    return out

"""Stage (c): lift-splat BEV pooling module backed by ``libmuvo_b200.so``.

Drop-in for ``muvo/models/frustum_pooling.py`` (``FrustumPooling``, ``QuickCumsum``,
``cumsum_trick``, ``quick_cumsum``, ``gen_dx_bx``) and ``VoxelsSumming``
(muvo/layers/layers.py:326-357).  Same constructor, buffers (``bev_intrinsics`` is the
only persistent one, so released checkpoints load with ``strict=True``), call
signatures and output layout ``(B, C*nz, ny, nx)``.

What stays in torch (plumbing): the frustum grid and the camera->ego->BEV geometry, using
the reference's own op sequence so the integer cell ids are bit-identical on the same
device (the kernel only ever sees integer cell ids).  What moves into CUDA: everything
that touches the feature tensor -- no reshape copy, no boolean compaction, no sort of
feature rows, no prefix sum; x is read once in place through its strides.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib, _ws

_FLOAT_DTYPES = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def bev_params_to_intrinsics(size, scale, offsetx):
    """muvo/utils/geometry_utils.py:8-19."""
    return np.array([[1 / scale, 0, size[0] / 2 + offsetx],
                     [0, -1 / scale, size[1] / 2],
                     [0, 0, 1]], dtype=np.float32)


def intrinsics_inverse(intrinsics):
    """muvo/utils/geometry_utils.py:22-34 (closed form)."""
    fx, fy = intrinsics[..., 0, 0], intrinsics[..., 1, 1]
    cx, cy = intrinsics[..., 0, 2], intrinsics[..., 1, 2]
    one, zero = torch.ones_like(fx), torch.zeros_like(fx)
    return torch.stack((torch.stack((1 / fx, zero, -cx / fx), -1),
                        torch.stack((zero, 1 / fy, -cy / fy), -1),
                        torch.stack((zero, zero, one), -1)), -2)


def gen_dx_bx(size, scale, offsetx):
    """muvo/models/frustum_pooling.py:10-20."""
    xbound = [-size[0] * scale / 2 - offsetx * scale, size[0] * scale / 2 - offsetx * scale, scale]
    ybound = [-size[1] * scale / 2, size[1] * scale / 2, scale]
    zbound = [-10.0, 10.0, 20.0]
    rows = [xbound, ybound, zbound]
    dx = torch.Tensor([row[2] for row in rows])
    bx = torch.Tensor([row[0] + row[2] / 2.0 for row in rows])
    nx = torch.LongTensor([np.round((row[1] - row[0]) / row[2]) for row in rows])
    return dx, bx, nx


# --------------------------------------------------------------------------- kernels as autograd functions
def _strides_bpc(x: torch.Tensor):
    """View ``x (B,N,D,H,W,C)`` as ``[B, n_pts, C]`` with element strides, without copying when possible."""
    B, N, D, H, W, Cc = x.shape
    n_pts = N * D * H * W
    s = x.stride()
    # the point index p = ((n*D + d)*H + h)*W + w must be expressible with ONE stride
    sp = s[4]
    ok = True
    ext = W
    for dim in (3, 2, 1):
        if x.shape[dim] != 1 and s[dim] != sp * ext:
            ok = False
            break
        ext *= x.shape[dim]
    if not ok or (B > 1 and s[0] < 0) or sp <= 0 or s[5] <= 0:
        x = x.contiguous()
        s = x.stride()
        sp = s[4]
    return x, int(s[0]), int(sp), int(s[5]), n_pts


class _BevPool(torch.autograd.Function):
    """out[b, c, cell] = sum over points p of frame b with cell[b,p] == cell of x[b,p,c] (ascending p)."""

    @staticmethod
    def forward(ctx, x, cell, n_cells):
        _lib.require_cuda(x, cell)
        if x.dtype not in _FLOAT_DTYPES:
            raise TypeError(f"unsupported feature dtype {x.dtype}")
        lib = _lib.load()
        xv, sb, sp, sc, n_pts = _strides_bpc(x)
        B, Cc = x.shape[0], x.shape[5]
        dev = x.device
        cell = cell.reshape(B, n_pts).contiguous()
        out = torch.empty((B, Cc, n_cells), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = _lib.current_stream(dev)
            nb = C.c_size_t(0)
            _lib.check(lib.muvo_bev_pool_workspace_bytes(B, n_pts, n_cells, C.byref(nb)), "muvo_bev_pool_workspace_bytes")
            ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
            rc = lib.muvo_bev_pool_fwd(_lib.ptr(xv), _FLOAT_DTYPES[xv.dtype], sb, sp, sc, _lib.ptr(cell), B, n_pts, Cc,
                                       n_cells, out.data_ptr(), ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "muvo_bev_pool_fwd")
        ctx.save_for_backward(cell)
        ctx.meta = (tuple(x.shape), x.dtype, n_cells, x.stride())
        ctx.fwd_ws = ws if lib.muvo_bev_pool_is_streamed(xv.element_size(), _lib.ptr(xv), sb, sp, sc, B, n_pts, Cc, n_cells) else None
        ctx.mark_non_differentiable(cell)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (cell,) = ctx.saved_tensors
        shape, dtype, n_cells, xstride = ctx.meta
        return bev_pool_backward(grad_out, cell, shape, dtype, n_cells, xstride, ctx.fwd_ws), None, None


class _BevPoolMasked(torch.autograd.Function):
    """``bev_pool(x, fold_mask(cell0, mask))`` as ONE library call (``muvo_bev_pool_fwd_masked``): the mask is folded in by
    the kernel that builds the per-chunk cell lists, and (B, C, D, H, W) memory is streamed instead of gathered."""

    @staticmethod
    def forward(ctx, x, cell0, mask, n_cells, plan=None):
        _lib.require_cuda(x, cell0)
        if x.dtype not in _FLOAT_DTYPES:
            raise TypeError(f"unsupported feature dtype {x.dtype}")
        lib = _lib.load()
        xv, sb, sp, sc, n_pts = _strides_bpc(x)
        B, Cc = x.shape[0], x.shape[5]
        dev = x.device
        cell0 = cell0.reshape(B, n_pts).contiguous()
        m = None
        if mask is not None and mask.numel() > 0:
            m = mask.reshape(-1)
            m = (m if m.dtype in (torch.bool, torch.uint8) else m.bool()).contiguous().view(torch.uint8)
            if m.numel() != cell0.numel():
                raise ValueError("mask and cell ids differ in size")
        cell = torch.empty_like(cell0)
        out = torch.empty((B, Cc, n_cells), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = _lib.current_stream(dev)
            nb = C.c_size_t(0)
            _lib.check(lib.muvo_bev_pool_workspace_bytes(B, n_pts, n_cells, C.byref(nb)), "muvo_bev_pool_workspace_bytes")
            ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
            rc = lib.muvo_bev_pool_fwd_masked(_lib.ptr(xv), _FLOAT_DTYPES[xv.dtype], sb, sp, sc, _lib.ptr(cell0),
                                              _lib.ptr(m) if m is not None else None, _lib.ptr(cell),
                                              plan.data_ptr() if plan is not None else None, B, n_pts, Cc, n_cells,
                                              out.data_ptr(), ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "muvo_bev_pool_fwd_masked")
        ctx.save_for_backward(cell)
        ctx.meta = (tuple(x.shape), x.dtype, n_cells, x.stride())
        ctx.fwd_ws = ws if lib.muvo_bev_pool_is_streamed(xv.element_size(), _lib.ptr(xv), sb, sp, sc, B, n_pts, Cc, n_cells) else None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (cell,) = ctx.saved_tensors
        shape, dtype, n_cells, xstride = ctx.meta
        return bev_pool_backward(grad_out, cell, shape, dtype, n_cells, xstride, ctx.fwd_ws), None, None, None, None


def build_plan(cell0: torch.Tensor, n_cells: int) -> torch.Tensor:
    """Mask-independent per-chunk sorted cell lists for ``bev_pool_masked(..., plan=)`` (``muvo_bev_plan_build``); valid for
    this exact ``cell0 (B, n_pts) int32`` only."""
    _lib.require_cuda(cell0)
    lib = _lib.load()
    B, n_pts = cell0.shape
    nb = C.c_size_t(0)
    _lib.check(lib.muvo_bev_plan_bytes(B, n_pts, C.byref(nb)), "muvo_bev_plan_bytes")
    plan = torch.empty(nb.value, dtype=torch.uint8, device=cell0.device)
    c0 = cell0.contiguous()
    with torch.cuda.device(cell0.device):
        rc = lib.muvo_bev_plan_build(_lib.ptr(c0), B, n_pts, int(n_cells), plan.data_ptr(), plan.numel(), _lib.current_stream(cell0.device))
    _lib.check(rc, "muvo_bev_plan_build")
    return plan


def bev_pool_backward(grad_out, cell, shape, dtype, n_cells, xstride=None, fwd_ws=None):
    """``grad_x[b,p,c] = cell[b,p] >= 0 ? grad_out[b,c,cell[b,p]] : 0`` (frustum_pooling.py:52-60 + index backward).

    The gradient is written in the producer's memory format: if x was a permuted ``(B,C,D,H,W)`` view
    (mile.py:517-521) so is grad_x, and the permute / mul backward consume it without a copy.
    """
    B, N, D, H, W, Cc = shape
    n_pts = N * D * H * W
    dev = grad_out.device
    go = grad_out.contiguous().float()
    if xstride is None or (xstride[5] >= xstride[4] and xstride[4] == 1):
        base = torch.empty((B, Cc, N, D, H, W), dtype=dtype, device=dev)
        gx = base.permute(0, 2, 3, 4, 5, 1)
    else:
        gx = torch.empty(shape, dtype=dtype, device=dev)
    gv, sb, sp, sc, _ = _strides_bpc(gx)
    with torch.cuda.device(dev):
        # fwd_ws: the workspace of a STREAMED forward (its per-chunk cell lists are reused: tiles built in shared memory,
        # written with TMA bulk stores); otherwise the gather kernel
        rc = _lib.load().muvo_bev_pool_bwd_streamed(go.data_ptr(), _lib.ptr(cell), B, n_pts, Cc, n_cells, gv.data_ptr(),
                                                    _FLOAT_DTYPES[dtype], sb, sp, sc,
                                                    fwd_ws.data_ptr() if fwd_ws is not None else None,
                                                    fwd_ws.numel() if fwd_ws is not None else 0, _lib.current_stream(dev))
    _lib.check(rc, "muvo_bev_pool_bwd_streamed")
    return gx


def bev_pool(x: torch.Tensor, cell: torch.Tensor, n_cells: int) -> torch.Tensor:
    """``x (B,N,D,H,W,C)`` any strides, ``cell (B, N*D*H*W) int32`` (-1 = dropped) -> ``(B, C, n_cells)`` fp32."""
    n_cells = int(n_cells)
    if n_cells > max_cells_per_pass():
        return _pool_in_windows(x, cell, n_cells)
    return _BevPool.apply(x, cell, n_cells)


def max_cells_per_pass() -> int:
    """Cells one library call handles (the index sort keeps one histogram per warp in shared memory: 12 800).  MUVO's
    shipped configs use 48*48*1 = 2304; bigger grids (BEV.SIZE 192x192 at FEATURE_DOWNSAMPLE 1, nz > 1, ...) are pooled in
    windows of that many cells, see ``_pool_in_windows``."""
    return int(_lib.load().muvo_bev_pool_max_cells())


def _pool_in_windows(x, cell, n_cells):
    """Grids above ``max_cells_per_pass()``: the cell axis is cut into windows [w0, w1) and every window is pooled by its own
    library call on ``cell - w0`` (points of other windows dropped), i.e. the same kernels, the same per-cell summation order and
    the same gradients (autograd sums the windows' disjoint contributions to grad_x); the cell ids are re-based with three torch
    elementwise ops per window.  The reference handles any grid (frustum_pooling.py:131-187); so does this."""
    step = max_cells_per_pass()
    B = x.shape[0]
    cell = cell.reshape(B, -1)
    outs = []
    for w0 in range(0, n_cells, step):
        w1 = min(w0 + step, n_cells)
        cw = torch.where((cell >= w0) & (cell < w1), cell - w0, torch.full_like(cell, -1))
        outs.append(_BevPool.apply(x, cw, w1 - w0))
    return torch.cat(outs, dim=2)


def bev_pool_masked(x: torch.Tensor, cell0: torch.Tensor, mask, n_cells: int, plan=None) -> torch.Tensor:
    """``x (B,N,D,H,W,C)``, mask-independent ``cell0 (B, n_pts) int32``, ``mask`` bool/uint8 with n_pts entries per frame (or
    empty / None) -> ``(B, C, n_cells)`` fp32: what ``bev_pool(x, fold_mask(cell0, mask), n_cells)`` returns.  ``plan`` =
    ``build_plan(cell0, n_cells)`` (optional, cached by the caller) removes the per-call sort."""
    n_cells = int(n_cells)
    if n_cells > max_cells_per_pass():
        has_mask = mask is not None and mask.numel() > 0
        return _pool_in_windows(x, fold_mask(cell0, mask) if has_mask else cell0, n_cells)
    return _BevPoolMasked.apply(x, cell0, mask, n_cells, plan)


def fold_mask(cell0: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """``cell0`` int32 with ``mask`` (bool / uint8, same number of elements) applied: dropped points become -1."""
    _lib.require_cuda(cell0, mask)
    m = mask.reshape(-1)
    m = (m if m.dtype in (torch.bool, torch.uint8) else m.bool()).contiguous().view(torch.uint8)
    c0 = cell0.contiguous()
    if m.numel() != c0.numel():
        raise ValueError("mask and cell ids differ in size")
    out = torch.empty_like(c0)
    with torch.cuda.device(c0.device):
        rc = _lib.load().muvo_bev_fold_mask(_lib.ptr(c0), _lib.ptr(m), c0.numel(), _lib.ptr(out), _lib.current_stream(c0.device))
    _lib.check(rc, "muvo_bev_fold_mask")
    return out


class _LiftSplat(torch.autograd.Function):
    """Fused lift-splat: ``out[b,c,cell] = sum_p depth[b,p] * feat[b,c,hw(p)]`` over the kept frustum
    points of the cell -- what ``bev_pool`` returns for the lifted tensor of mile.py:517-521, without building it."""

    @staticmethod
    def forward(ctx, feat, depth, cell, n_cells, plan=None, mask=None):
        _lib.require_cuda(feat, depth, cell)
        lib = _lib.load()
        B, Cc, H, W = feat.shape
        D = depth.shape[1]
        if depth.shape != (B, D, H, W):
            raise ValueError("depth must be (B, D, H, W) matching feat (B, C, H, W)")
        dev = feat.device
        feat_cl = feat.detach().float().permute(0, 2, 3, 1).contiguous()          # [B, HW, C]: one C-vector per pixel
        dep = depth.detach().float().contiguous()
        cell = cell.reshape(B, D * H * W).contiguous()
        out = torch.empty((B, Cc, n_cells), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = _lib.current_stream(dev)
            nb = C.c_size_t(0)
            _lib.check(lib.muvo_bev_pool_workspace_bytes(B, D * H * W, n_cells, C.byref(nb)), "muvo_bev_pool_workspace_bytes")
            ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
            if plan is not None:                 # cached mask-independent sort: only the mask filter runs per call
                rc = lib.muvo_lift_splat_fwd_planned(feat_cl.data_ptr(), dep.data_ptr(), plan.data_ptr(), plan.numel(), _lib.ptr(mask), B, D,
                                                     H * W, Cc, n_cells, out.data_ptr(), ws.data_ptr(), ws.numel(), stream)
            else:
                rc = lib.muvo_lift_splat_fwd(feat_cl.data_ptr(), dep.data_ptr(), _lib.ptr(cell), B, D, H * W, Cc, n_cells,
                                             out.data_ptr(), ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "muvo_lift_splat_fwd")
        ctx.save_for_backward(feat_cl, dep, cell)
        ctx.meta = (B, Cc, D, H, W, n_cells, feat.dtype, depth.dtype)
        ctx.mark_non_differentiable(cell)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        feat_cl, dep, cell = ctx.saved_tensors
        B, Cc, D, H, W, n_cells, fdt, ddt = ctx.meta
        dev = grad_out.device
        gout_cl = grad_out.float().permute(0, 2, 1).contiguous()                    # [B, n_cells, C]
        gdepth = torch.empty((B, D, H, W), dtype=torch.float32, device=dev)
        gfeat_cl = torch.empty((B, H, W, Cc), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load().muvo_lift_splat_bwd(gout_cl.data_ptr(), feat_cl.data_ptr(), dep.data_ptr(), _lib.ptr(cell), B, D,
                                                 H * W, Cc, n_cells, gdepth.data_ptr(), gfeat_cl.data_ptr(),
                                                 _lib.current_stream(dev))
        _lib.check(rc, "muvo_lift_splat_bwd")
        return gfeat_cl.permute(0, 3, 1, 2).to(fdt), gdepth.to(ddt), None, None, None, None


def lift_splat(feat: torch.Tensor, depth: torch.Tensor, cell: torch.Tensor, n_cells: int, plan=None, mask=None) -> torch.Tensor:
    """``feat (B,C,H,W)``, ``depth (B,D,H,W)``, ``cell (B, D*H*W) int32`` (-1 = dropped) -> ``(B, C, n_cells)`` fp32.
    With ``plan`` (:func:`build_lift_splat_plan` of the mask-independent cell ids) and the ``mask`` that was folded into
    ``cell``, the forward skips the per-call cell sort (``cell`` is still what the backward walks)."""
    if int(n_cells) > max_cells_per_pass():
        raise ValueError(f"the fused lift-splat handles at most {max_cells_per_pass()} BEV cells (nx*ny*nz), got {n_cells}; "
                         "FrustumPooling.forward / bev_pool pool bigger grids in windows")
    if plan is not None:
        m = None
        if mask is not None and mask.numel() > 0:
            m = mask.reshape(-1)
            m = (m if m.dtype in (torch.bool, torch.uint8) else m.bool()).contiguous().view(torch.uint8)
            if m.numel() != cell.numel():
                raise ValueError("mask and cell ids differ in size")
        return _LiftSplat.apply(feat, depth, cell, int(n_cells), plan, m)
    return _LiftSplat.apply(feat, depth, cell, int(n_cells))


def build_lift_splat_plan(cell0: torch.Tensor, n_cells: int) -> torch.Tensor:
    """The mask-independent cell sort of the fused lift-splat (``muvo_lift_splat_plan_build``), to be cached per camera rig:
    ``cell0 (B, n_pts) int32`` WITHOUT the mask folded in.  Pass it to :func:`lift_splat` together with the mask."""
    _lib.require_cuda(cell0)
    lib = _lib.load()
    B, n_pts = cell0.shape
    dev = cell0.device
    nb, nw = C.c_size_t(0), C.c_size_t(0)
    _lib.check(lib.muvo_lift_splat_plan_bytes(B, n_pts, int(n_cells), C.byref(nb)), "muvo_lift_splat_plan_bytes")
    _lib.check(lib.muvo_bev_pool_workspace_bytes(B, n_pts, int(n_cells), C.byref(nw)), "muvo_bev_pool_workspace_bytes")
    plan = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    ws = torch.empty(nw.value, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.muvo_lift_splat_plan_build(_lib.ptr(cell0.contiguous()), B, n_pts, int(n_cells), plan.data_ptr(), plan.numel(), ws.data_ptr(),
                                            ws.numel(), _lib.current_stream(dev))
    _lib.check(rc, "muvo_lift_splat_plan_build")
    return plan


class QuickCumsum(torch.autograd.Function):
    """Sorted-rank segment sum; drop-in for ``QuickCumsum`` (frustum_pooling.py:34-60)."""

    @staticmethod
    def forward(ctx, x, geom_feats, ranks):
        _lib.require_cuda(x, ranks)
        lib = _lib.load()
        n = int(x.shape[0])
        Cc = int(x[0].numel()) if n else int(np.prod(x.shape[1:]))
        dev = x.device
        in_dtype = x.dtype
        xf = x.reshape(n, Cc).contiguous().float()                 # cumsum runs in fp32 under autocast too
        rk = ranks.contiguous().to(torch.int64)
        seg_id = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
        last_row = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
        n_seg_t = torch.zeros((1,), dtype=torch.int32, device=dev)
        x_seg = torch.empty((max(n, 1), Cc), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = _lib.current_stream(dev)
            nb = C.c_size_t(0)
            _lib.check(lib.muvo_segment_sum_workspace_bytes(n, C.byref(nb)), "muvo_segment_sum_workspace_bytes")
            ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
            rc = lib.muvo_segment_sum_fwd(_lib.ptr(xf), _lib.ptr(rk), n, Cc, seg_id.data_ptr(), n_seg_t.data_ptr(),
                                          x_seg.data_ptr(), last_row.data_ptr(), ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "muvo_segment_sum_fwd")
        n_seg = int(n_seg_t.item())          # data-dependent output shape: one sync, as the reference's x[kept]
        x_seg = x_seg[:n_seg].reshape((n_seg,) + tuple(x.shape[1:])).to(in_dtype)
        geom_out = geom_feats[last_row[:n_seg]]
        ctx.save_for_backward(seg_id[:n])
        ctx.meta = (n, Cc, tuple(x.shape), in_dtype)
        ctx.mark_non_differentiable(geom_out)
        return x_seg, geom_out

    @staticmethod
    def backward(ctx, gradx, gradgeom):
        (seg_id,) = ctx.saved_tensors
        n, Cc, shape, in_dtype = ctx.meta
        dev = gradx.device
        g = gradx.reshape(-1, Cc).contiguous().float()
        out = torch.empty((n, Cc), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load().muvo_segment_sum_bwd(_lib.ptr(g), _lib.ptr(seg_id), n, Cc, _lib.ptr(out), _lib.current_stream(dev))
        _lib.check(rc, "muvo_segment_sum_bwd")
        return out.reshape(shape).to(in_dtype), None, None


class VoxelsSumming(QuickCumsum):
    """Drop-in for ``muvo.layers.layers.VoxelsSumming`` (layers.py:326-357): same op, same signature."""


def quick_cumsum(x, geom_feats, ranks):
    return QuickCumsum.apply(x, geom_feats, ranks)


def cumsum_trick(x, geom_feats, ranks):
    """``cumsum_trick`` (frustum_pooling.py:23-31): same values as ``QuickCumsum`` and, like the reference's plain-torch
    version (used in training when ``use_quickcumsum=False``, :176-178), differentiable in ``x``."""
    return QuickCumsum.apply(x, geom_feats, ranks)


# --------------------------------------------------------------------------- the module
class FrustumPooling(nn.Module):
    def __init__(self, size, scale, offsetx, dbound, downsample, use_quickcumsum=True):
        """Pools camera frustums into Birds Eye View (constructor of frustum_pooling.py:68-95)."""
        super().__init__()
        self.register_buffer('bev_intrinsics', torch.tensor(bev_params_to_intrinsics(size, scale, offsetx)))
        dx, bx, nx = gen_dx_bx(size, scale, offsetx)
        self.nx_constant = nx.numpy().tolist()
        self.register_buffer('dx', dx, persistent=False)
        self.register_buffer('bx', bx, persistent=False)
        self.register_buffer('nx', nx, persistent=False)
        self.use_quickcumsum = use_quickcumsum
        self.dbound = dbound
        ds = torch.arange(self.dbound[0], self.dbound[1], self.dbound[2], dtype=torch.float32)
        self.D = len(ds)
        self.register_buffer('ds', ds, persistent=False)
        self.downsample = downsample
        self.register_buffer('frustum', torch.zeros(0, ), persistent=False)
        self._geom_cache = None          # (intrinsics, pose, frustum shape) -> mask-independent cell ids; not part of state_dict

    def initialize_frustum(self, image):
        """frustum_pooling.py:97-109."""
        if self.frustum.shape[0] == 0:
            device = image.device
            fH, fW = image.shape[-3:-1]
            ogfH, ogfW = fH * self.downsample, fW * self.downsample
            ds = self.ds.view(-1, 1, 1).expand(-1, fH, fW)
            xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float, device=device).view(1, 1, fW).expand(self.D, fH, fW)
            ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float, device=device).view(1, fH, 1).expand(self.D, fH, fW)
            self.frustum = torch.stack((xs, ys, ds), -1)

    def get_geometry(self, rots, trans, intrins):
        """frustum_pooling.py:111-129: (x,y,z) ego-frame locations, ``B x N x D x H x W x 3``."""
        B, N = trans.shape[:2]
        points = self.frustum.unsqueeze(0).unsqueeze(0).unsqueeze(-1)
        points = torch.cat((points[:, :, :, :, :, :2] * points[:, :, :, :, :, 2:3], points[:, :, :, :, :, 2:3]), 5)
        combine = rots.matmul(intrinsics_inverse(intrins))
        points = combine.view(B, N, 1, 1, 1, 3, 3).matmul(points).squeeze(-1)
        points += trans.view(B, N, 1, 1, 1, 3)
        return points

    def cell_ids(self, geom_feats, mask):
        """Integer BEV cell per frustum point, -1 where dropped (frustum_pooling.py:139-163), int32 ``(B, n_pts)``."""
        B = geom_feats.shape[0]
        g = geom_feats.reshape(-1, 3)
        gx = (g[:, 0] * self.bev_intrinsics[0, 0] + self.bev_intrinsics[0, 2]).long()        # :142,:146 (.long() truncates)
        gy = (g[:, 1] * self.bev_intrinsics[1, 1] + self.bev_intrinsics[1, 2]).long()        # :143
        gz = ((g[:, 2] - self.bx[2] + self.dx[2] / 2.) / self.dx[2]).long()                  # :145
        nx, ny, nz = self.nx_constant
        kept = (gx >= 0) & (gx < nx) & (gy >= 0) & (gy < ny) & (gz >= 0) & (gz < nz)         # :159-161
        if len(mask) > 0:
            kept = kept & mask.reshape(-1).bool()                                            # :153-156
        cell = (gz * ny + gy) * nx + gx
        cell = torch.where(kept, cell, torch.full_like(cell, -1)).to(torch.int32)
        return cell.view(B, -1)

    def voxel_pooling(self, geom_feats, x, mask):
        """frustum_pooling.py:131-187 -> ``(B, C*nz, ny, nx)``."""
        B, N, D, H, W, Cc = x.shape
        nx, ny, nz = self.nx_constant
        cell = self.cell_ids(geom_feats, mask)
        out = bev_pool(x, cell, nx * ny * nz)                 # (B, C, nz*ny*nx), fp32 (cumsum's autocast dtype)
        # (B, C, nz, ny, nx) -> cat(unbind(dim=2), 1) == (B, nz*C, ny, nx) with z-major channel blocks (:185)
        out = out.view(B, Cc, nz, ny, nx)
        if nz == 1:
            return out.view(B, Cc, ny, nx)
        return out.permute(0, 2, 1, 3, 4).reshape(B, nz * Cc, ny, nx)

    def cached_cell_ids(self, intrinsics, pose, frustum_like):
        """Mask-independent cell id of every frustum point, cached.

        The geometry (frustum_pooling.py:111-129) and the integer cell of a frustum point (:139-163 without the mask)
        depend only on the intrinsics, the extrinsics and the frustum shape -- constant for a fixed camera rig -- so they
        are computed once with the reference's own torch ops (bit-identical cell ids) and reused; per call only the top-k
        mask is folded in (one small kernel).  Cache hit = the same tensor objects (unchanged in place), or equal values
        (``torch.equal``: two tiny kernels and one sync, still far below the ~25 eager launches + batched gemv it replaces;
        the reference's own ``x[kept]`` indexing synchronises every call as well)."""
        self.initialize_frustum(frustum_like)
        c = self._geom_cache
        key = (tuple(intrinsics.shape), tuple(pose.shape), tuple(self.frustum.shape), intrinsics.device, intrinsics.dtype)
        if c is not None and c["key"] == key:
            same_obj = (c["K_id"] == (id(intrinsics), intrinsics._version) and c["E_id"] == (id(pose), pose._version))
            if same_obj or (torch.equal(intrinsics, c["K"]) and torch.equal(pose, c["E"])):
                c["K_id"], c["E_id"] = (id(intrinsics), intrinsics._version), (id(pose), pose._version)
                c["hits"] += 1
                return c["cell0"]
        with torch.no_grad():
            geom = self.get_geometry(pose[..., :3, :3], pose[..., :3, 3:], intrinsics)
            cell0 = self.cell_ids(geom, torch.zeros(0))
        self._geom_cache = {"key": key, "K": intrinsics.detach().clone(), "E": pose.detach().clone(), "cell0": cell0, "hits": 0,
                            "K_id": (id(intrinsics), intrinsics._version), "E_id": (id(pose), pose._version)}
        return cell0

    def _pool_cells(self, x, cell0, mask):
        B, N, D, H, W, Cc = x.shape
        nx, ny, nz = self.nx_constant
        c = self._geom_cache
        plan = None
        if c is not None and c["cell0"] is cell0 and cell0.shape[0] == B and nx * ny * nz <= max_cells_per_pass():
            if c.get("plan") is None:
                c["plan"] = build_plan(cell0, nx * ny * nz)
            plan = c["plan"]
        out = bev_pool_masked(x, cell0, mask if len(mask) > 0 else None, nx * ny * nz, plan).view(B, Cc, nz, ny, nx)
        return out.view(B, Cc, ny, nx) if nz == 1 else out.permute(0, 2, 1, 3, 4).reshape(B, nz * Cc, ny, nx)

    def forward(self, x, intrinsics, pose, mask=torch.zeros(0)):
        """frustum_pooling.py:189-209 (geometry + cell ids cached per camera, see :meth:`cached_cell_ids`)."""
        cell0 = self.cached_cell_ids(intrinsics, pose, x)
        return self._pool_cells(x, cell0, mask).type_as(x)

    def lift_splat(self, feat, depth, intrinsics, pose, mask=torch.zeros(0)):
        """Opt-in fused replacement of ``forward((depth[:,None] * feat[:,:,None])[:,None].permute(0,1,3,4,5,2), ...)``
        (muvo/models/mile.py:517-523): same output, the (B,C,D,H,W) outer product is never materialised and the
        gradients go straight to ``feat`` and ``depth``.  ``feat (B,C,fH,fW)``, ``depth (B,D,fH,fW)``, one camera."""
        B, Cc, H, W = feat.shape
        cell0 = self.cached_cell_ids(intrinsics, pose, feat.new_zeros((1, 1, 1, H, W, 1)))
        nx, ny, nz = self.nx_constant
        cell = fold_mask(cell0, mask) if len(mask) > 0 else cell0
        plan = None
        c = self._geom_cache
        if c is not None and c["cell0"] is cell0 and cell0.shape[0] == B and nx * ny * nz <= max_cells_per_pass():
            if c.get("ls_plan") is None:
                c["ls_plan"] = build_lift_splat_plan(cell0.reshape(B, -1), nx * ny * nz)
            plan = c["ls_plan"]
        out = lift_splat(feat, depth, cell, nx * ny * nz, plan, mask if len(mask) > 0 else None).view(B, Cc, nz, ny, nx)
        out = out.view(B, Cc, ny, nx) if nz == 1 else out.permute(0, 2, 1, 3, 4).reshape(B, nz * Cc, ny, nx)
        return out.type_as(feat)

    def get_depth_map(self, depth):
        """frustum_pooling.py:211-217."""
        ds = self.ds.view(1, -1, 1, 1)
        depth = (ds * depth).sum(1, keepdim=True)
        depth = nn.functional.interpolate(depth, scale_factor=float(self.downsample), mode='bilinear', align_corners=False)
        return depth

"""rasterize(): drop-in for splat_py.rasterize.rasterize (splat_py/rasterize.py:18-112).

Same nine positional arguments, same return triple ``(image [H,W,3], culling_mask [N] bool,
uv [M,2])``, same values.  ``uv`` is a graph tensor whose ``.grad`` after ``backward()`` is the
render pass's gradient on the projected means, exactly what the reference trainer reads for
densification (splat_py/trainer.py:360,379-385).

Implementation: two autograd nodes around the native fused kernels

    _ProjectGaussians   params -> uv [M,2], carrier          (gsr_preprocess_forward + binning)
    _CompositeTiles     uv, carrier -> image                  (gsr_render_forward)

(The tile renderer's forward kernel is enqueued by the projection node, right behind the binning and BEFORE the
host reads the pair count; `_CompositeTiles` only ties its image into the graph.)

``carrier`` is an uninitialised [12N] tensor ([9N] with the record stream) that only carries gradient: the
render backward returns its per-Gaussian sums as the carrier's gradient — interleaved rows
``rgb3 opacity | uv2 conic0 conic1 | conic2 pad3`` (planar rgb | opacity | uv | conic with the record stream) — and
the per-Gaussian backward consumes them.  One host read per forward (the pair count, which sizes the pair
buffers) and it is taken off the critical path: the buffers are sized from the previous views' counts and the
count is only checked afterwards; the reference path has ~25 syncs.

``use_sh_precompute=False`` (per-pixel view directions) is routed through the operator-by-operator
path `rasterize_unfused`, which mirrors the reference's structure on this library's operators.
"""
from __future__ import annotations

import weakref

import torch

from . import native
from .cuda_autograd_functions import (
    CameraPointProjection,
    ComputeConic,
    ComputeProjectionJacobian,
    ComputeSigmaWorld,
    PrecomputeRGBFromSH,
    RenderImage,
)
from .structs import Tiles
from .tile_culling import get_splats
from .utils import compute_rays_in_world_frame, transform_points_torch


import os

# world->camera inside the fused kernel (saves a 2.3 ms torch.matmul at 3M points).  The kernel reproduces the
# rounding order of the cuBLAS kernel torch dispatches to for large batches (verified bitwise on B200 for
# N in {2000 ... 3M}); below IN_KERNEL_TRANSFORM_MIN_N, or with GSR_TORCH_TRANSFORM=1, positions are formed by
# torch.matmul exactly as the reference does (splat_py/utils.py:60-72).
IN_KERNEL_TRANSFORM = os.environ.get("GSR_TORCH_TRANSFORM", "0") != "1"
# the forward records which (8x4 pixel block, splat) pairs contributed; the backward then visits exactly those
# (GSR_NO_MASKS=1: the backward redoes the forward's conservative footprint test instead)
USE_CONTRIBUTION_MASKS = os.environ.get("GSR_NO_MASKS", "0") != "1"
# the tile kernels fetch their records through the sorted pair list themselves (cp.async gather); with
# GSR_RECORD_STREAM=1 a separate kernel first writes the depth-sorted, tile-contiguous record stream they then read
# with bulk copies (the round-1 arrangement)
USE_RECORD_GATHER = os.environ.get("GSR_RECORD_STREAM", "0") != "1"
# pair emission expands the tile-hit masks of the per-gaussian stage (GSR_RETEST_TILES=1: repeats the OBB tests)
USE_TILE_MASKS = os.environ.get("GSR_RETEST_TILES", "0") != "1"
IN_KERNEL_TRANSFORM_MIN_N = 16384
CUBLAS_BATCH_CHUNK = 65535   # gridDim limit cuBLAS batches against
SMALL_BATCH = 1024           # chunks below ~300 use another kernel (measured: 100 differs, 300 matches)


_TRANSFORM_OK = {}  # device index -> bool: did the in-kernel transform reproduce torch.matmul on this device?


def in_kernel_transform_ok(device, n: int = 70_000, seed: int = 1234) -> bool:
    """One-time self-check per device (a cuBLAS / torch upgrade could change the rounding order the fused
    kernel mimics): `n` random points (more than one 65535-matrix cuBLAS chunk) under a random rigid pose, camera
    z and uv from the in-kernel transform compared BIT FOR BIT with the reference's op
    (splat_py/utils.py:60-72, torch.matmul).  On a mismatch the fused path forms positions with torch.matmul
    itself from then on (the reference's own op on the GPU, +2.3 ms at 3M points) and says so once."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    ok = _TRANSFORM_OK.get(idx)
    if ok is not None:
        return ok
    import math
    import warnings

    ext = native()
    gen = torch.Generator().manual_seed(seed)
    xyz = (torch.rand(n, 3, generator=gen) * torch.tensor([8.0, 6.0, 10.0]) + torch.tensor([-4.0, -3.0, 1.0])).to(device)
    q = torch.randn(4, generator=gen)
    w, x, y, z = (q / q.norm()).tolist()
    T = torch.eye(4)
    T[:3, :3] = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                              [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                              [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T[:3, 3] = torch.randn(3, generator=gen)
    T = T.to(device)
    K = torch.tensor([[1200.0, 0.0, 960.0], [0.0, 1200.0, 540.0], [0.0, 0.0, 1.0]], device=device)
    quat = torch.zeros(n, 4, device=device)
    quat[:, 0] = 1.0
    zeros3, zeros1 = torch.zeros(n, 3, device=device), torch.zeros(n, device=device)
    big = 1e30
    with torch.no_grad():
        rec, zkey, _, _ = ext.fused_preprocess_forward(xyz, None, quat, zeros3, zeros1, zeros3, None, T, K, None, 1080,
                                                       1920, -big, big, big, 3.0, 0)
        ref = transform_points_torch(xyz, T)
        zref = ref[:, 2].contiguous().view(torch.int32)
        zref = torch.where(zref < 0, ~zref, zref | -2**31)
        uv = torch.zeros(n, 2, device=device)
        ext.camera_projection_cuda(ref, K, uv)
        fin = torch.isfinite(uv).all(dim=1) & torch.isfinite(rec[:, :2]).all(dim=1)
        same = bool((zkey == zref).all()) and bool((rec[:, 0:2][fin].contiguous().view(torch.int32)
                                                    == uv[fin].contiguous().view(torch.int32)).all())
    _TRANSFORM_OK[idx] = same
    if not same:
        warnings.warn("gaussian_splatting_b200: the in-kernel world->camera transform no longer reproduces "
                      "torch.matmul bit for bit on this device (cuBLAS changed its rounding order?); "
                      "falling back to torch.matmul for camera-frame positions", RuntimeWarning)
    return same


# ---- speculative sizing of the pair buffers (see _ProjectGaussians.forward) -----------------------------------------
SPECULATIVE_BINNING = os.environ.get("GSR_EAGER_COUNTS", "0") != "1"
_PAIR_HISTORY = {}   # (device index, N, H, W) -> pair counts of the most recent views
_PINNED = {}         # device index -> [pinned int64 ring, next slot]
PAIR_HEADROOM = 1.04  # capacity = max(recent P) * headroom + 64 Ki pairs


def _pair_capacity(key):
    hist = _PAIR_HISTORY.get(key)
    if not hist:
        return None
    cap = int(max(hist) * PAIR_HEADROOM) + 65536
    return min((cap + 127) // 128 * 128, 2**31 - 128)


def _note_pairs(key, P):
    hist = _PAIR_HISTORY.setdefault(key, [])
    hist.append(int(P))
    del hist[:-8]
    if len(_PAIR_HISTORY) > 64:  # scenes come and go (densification changes N): keep the table small
        for k in list(_PAIR_HISTORY)[:-32]:
            del _PAIR_HISTORY[k]


_SIDE = {}           # device index -> side stream for the count read-back


def _side_stream(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _SIDE.get(idx)
    if st is None:
        st = _SIDE[idx] = torch.cuda.Stream(device=device)
    return st


def _pinned_slot(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    ring = _PINNED.get(idx)
    if ring is None:
        ring = _PINNED[idx] = [torch.empty(16, dtype=torch.int64).pin_memory(), 0]
    ring[1] = (ring[1] + 1) % 16
    return ring[0][ring[1]:ring[1] + 1]


_CENTRE_OK = {}  # device index -> bool: does gsr_camera_centre reproduce torch's inverse on this device?


def native_centre_ok(device, n: int = 8, seed: int = 4321) -> bool:
    """One-time self-check per device: the camera centre from the library's single kernel (the arithmetic of
    cuSOLVER getrf + cuBLAS trsm as pinned on a B200, csrc/gsr_preprocess.cu `camera_centre`) against
    torch.linalg.inv_ex, BIT FOR BIT, on `n` random rigid poses and a few general matrices.  A cuSOLVER / cuBLAS
    upgrade that changes the rounding order makes rasterize() use torch's own inverse again (15 micro-kernels)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    ok = _CENTRE_OK.get(idx)
    if ok is not None:
        return ok
    import warnings

    ext = native()
    gen = torch.Generator().manual_seed(seed)
    same = True
    with torch.no_grad():
        for k in range(n):
            if k % 4 == 3:
                T = torch.eye(4)
                T[:3, :] = torch.randn(3, 4, generator=gen)
            else:
                q = torch.randn(4, generator=gen)
                w, x, y, z = (q / q.norm()).tolist()
                T = torch.eye(4)
                T[:3, :3] = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
                T[:3, 3] = torch.randn(3, generator=gen) * 3.0
            T = T.to(device)
            a = ext.camera_centre(T)
            b = torch.linalg.inv_ex(T)[0][:3, 3].contiguous()
            same = same and bool((a.view(torch.int32) == b.view(torch.int32)).all())
    _CENTRE_OK[idx] = same
    if not same:
        warnings.warn("gaussian_splatting_b200: gsr_camera_centre no longer reproduces torch.linalg.inv_ex bit for "
                      "bit on this device; using torch's inverse for the camera centre", RuntimeWarning)
    return same


import functools


@functools.lru_cache(maxsize=64)
def _depth_key_params(near, far):
    """(base, bits): depth keys are float bits of z minus `base`; `bits` low bits are significant."""
    import math
    import struct

    def fbits(v):
        return struct.unpack("<I", struct.pack("<f", v))[0]

    FLT_MAX = 3.4028234663852886e38
    if near > 0.0 and math.isfinite(far) and near < far <= FLT_MAX:
        b0, b1 = fbits(near), fbits(far)  # struct.pack rounds to nearest: a tiny `near` can round to 0.0f
        if b0 > 0 and b1 > b0:
            return b0, max(1, (b1 - b0).bit_length())
    # no usable positive range (near <= 0 or rounding to 0.0f, far infinite / beyond FLT_MAX / NaN): full 32-bit
    # order-preserving keys (depth_key() sets bit 31, so depth_bits MUST be 32 whenever the base is 0)
    return 0, 32


class _ViewState:
    """Non-differentiable per-view buffers shared by the two autograd nodes."""

    __slots__ = ("N", "M", "P", "H", "W", "visible", "vis_idx", "_ids_sorted", "ranges", "stream_rec", "gather",
                 "records", "keys_sorted", "id_bits", "slab_width",
                 "n_per_pixel", "w_per_pixel", "background", "profile", "grad_flat", "grad_out", "vis_idx32", "scan", "masks", "uv_ref", "uv_grad_emitted", "image", "speculation_overflowed")

    @property
    def ids_sorted(self):
        """int32 [P]: gaussian id of every sorted (gaussian, tile) pair (the reference's
        `sorted_gaussian_idx_by_splat_idx`).  When the id rides in the sort key and the tile kernels gather through
        the keys, nobody writes this array: it is extracted on demand."""
        if self._ids_sorted is not None and self._ids_sorted.numel() == 0 and self.P > 0 and self.keys_sorted is not None:
            mask = (1 << self.id_bits) - 1
            self._ids_sorted = (self.keys_sorted[:self.P] & mask).to(torch.int32)
        return self._ids_sorted

    def __init__(self, profile=None):
        self.profile = profile  # optional list: (stage name, start event, end event) per native call
        self.grad_flat = None   # set by the backward pass: flat buffer holding all parameter gradients
        self.grad_out = None    # optional caller-owned buffer (same layout) the backward writes them into
        self.uv_ref = None      # weak reference to the uv tensor handed to the caller
        self.uv_grad_emitted = False
        self.background = None  # set by rasterize() before the projection node runs (it pre-launches the renderer)
        self.image = None
        self.speculation_overflowed = False


class _stage:
    """CUDA-event bracket around one native stage (only when the caller asked for a profile)."""

    def __init__(self, state, name):
        self.rec = state.profile
        self.name = name

    def __enter__(self):
        if self.rec is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if self.rec is not None:
            self.e1.record()
            self.rec.append((self.name, self.e0, self.e1))


class _ProjectGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, quaternion, scale, opacity, rgb, sh, camera_T_world, K, state, cfg):
        ext = native()
        H, W, near, far, pad, mh = cfg
        depth_base, depth_bits = _depth_key_params(near, far)
        opacity_flat = opacity.reshape(-1)
        # camera-frame positions: by default formed exactly like the reference does (torch.matmul,
        # splat_py/utils.py:60-72), because their bits decide tile membership and the 1/255 skip and the
        # rounding order inside cuBLAS is not ours to pin; IN_KERNEL_TRANSFORM folds it into the kernel.
        N = xyz.shape[0]
        if IN_KERNEL_TRANSFORM and N >= IN_KERNEL_TRANSFORM_MIN_N and in_kernel_transform_ok(xyz.device):
            # cuBLAS runs the batched product in chunks of 65535 matrices; a short last chunk goes through
            # its small-batch kernel (different rounding order), so that tail is formed by torch itself
            tail = N % CUBLAS_BATCH_CHUNK
            xyz_cam = transform_points_torch(xyz[N - tail:], camera_T_world) if 0 < tail < SMALL_BATCH else None
        else:
            xyz_cam = transform_points_torch(xyz, camera_T_world)
        # camera centre for the SH view direction = inverse(camera_T_world)[:3, 3] with the bits of the reference's
        # torch.inverse (splat_py/rasterize.py:91-93): ONE native kernel that reproduces torch's LU arithmetic
        # (verified per device, native_centre_ok), else torch's own inverse (inv_ex: no host sync)
        centre = None
        if sh is not None:
            if native_centre_ok(xyz.device):
                centre = ext.camera_centre(camera_T_world)
            else:
                centre = torch.linalg.inv_ex(camera_T_world)[0][:3, 3].contiguous()
        with _stage(state, "preprocess_fwd"):
            # tile_mask / tile_win: the tiles each gaussian hits, tested once here and only expanded by the pair emission
            records, zkey, visible, scan, tile_mask, tile_win = ext.fused_preprocess_forward_tiles(
                xyz, xyz_cam, quaternion, scale, opacity_flat, rgb, sh, camera_T_world, K, centre, H, W, near, far,
                pad, mh, depth_base)
        record_masks = USE_CONTRIBUTION_MASKS and any(ctx.needs_input_grad[:6])  # a backward pass can follow
        background = state.background

        def counts(total):
            M_, P_ = total >> 32, total & 0xFFFFFFFF
            # M and P share one u64 scan (M<<32 | P) and the native entry points index pairs with int32: a scene
            # with >= 2^31 (gaussian, tile) pairs must fail loudly, not wrap into an empty render
            if M_ > N or P_ >= 2**31:
                raise RuntimeError(f"rasterize: {P_} (gaussian, tile) pairs / {M_} visible of {N} gaussians exceed "
                                   "the int32 pair index of the native path (P must be < 2^31)")
            return M_, P_

        gather = USE_RECORD_GATHER

        def bin_and_render(M_, P_, speculative):
            with _stage(state, "bin_sort_gather"):
                binned = ext.fused_bin(records, zkey, visible, scan, M_, P_, H, W, mh, depth_bits, speculative, gather,
                                       tile_mask if USE_TILE_MASKS else None, tile_win if USE_TILE_MASKS else None)
            with _stage(state, "render_fwd"):
                rendered = ext.fused_render_forward(records if gather else binned[2], binned[1], background, H, W, P_,
                                                    record_masks, binned[5], binned[0], binned[6], gather)
            return binned, rendered

        # The host needs M and P (the shape of the returned uv; the size of the pair buffers).  Reading them right
        # here stalls the GPU for a launch round trip (35 us resident, ~100 us when the step's host->device copies
        # share the queue).  Instead the pair buffers are sized from the pair counts of the recent views of this
        # (N, H, W) plus headroom, binning AND the tile renderer are enqueued, and only then the counts — copied to
        # pinned memory right behind the scan — are read: by then they are long there.  Padding behind the real
        # pairs sorts to the end and is ignored (gsr_emit_*'s capacity).  If a view overflows the guess, the two
        # stages run again with the exact sizes (and the guess grows); the first view of a kind reads eagerly.
        key = (xyz.device.index, N, H, W)
        cap = _pair_capacity(key) if (SPECULATIVE_BINNING and N > 0) else None
        if cap is None:
            total = int(scan[-1].item()) if N > 0 else 0
            M, P = counts(total)
            binned, rendered = bin_and_render(M, P, False)
        else:
            # the 8-byte copy goes on a side stream: on the main stream the binning kernels would queue behind it,
            # and it is slow when the step's own host<->device traffic shares the copy engines (25 us, measured)
            slot = _pinned_slot(xyz.device)
            side = _side_stream(xyz.device)
            scanned = torch.cuda.Event()
            scanned.record()
            with torch.cuda.stream(side):
                side.wait_event(scanned)
                slot.copy_(scan[-1:], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(side)
            binned, rendered = bin_and_render(N, cap, True)
            ready.synchronize()
            M, P = counts(int(slot.item()))
            if P > cap:  # the guess was too small: nothing usable was produced, run the two stages again
                binned, rendered = bin_and_render(M, P, False)
                state.speculation_overflowed = True
            else:
                ids_c, ranges_c, stream_c, vis_c, uv_c, keys_c, idb = binned
                binned = (ids_c.narrow(0, 0, P) if ids_c.numel() else ids_c, ranges_c, stream_c, vis_c.narrow(0, 0, M),
                          uv_c.narrow(0, 0, M), keys_c, idb)
        _note_pairs(key, P)
        ids_sorted, ranges, stream_rec, vis_idx, uv, keys_sorted, id_bits = binned
        state.gather, state.records, state.keys_sorted, state.id_bits = gather, records, keys_sorted, int(id_bits)
        state.image, state.n_per_pixel, state.w_per_pixel, state.masks = rendered
        state.N, state.M, state.P, state.H, state.W = xyz.shape[0], M, P, H, W
        state.visible = visible
        state.vis_idx = state.vis_idx32 = vis_idx                # int32 [M]: visible gaussians, ascending
        state._ids_sorted, state.ranges, state.stream_rec = ids_sorted, ranges, stream_rec
        state.scan = scan
        # gradient slab rows: 12 floats per gaussian (interleaved, gather mode) or 9 (planar, record-stream mode)
        state.slab_width = 12 if gather else 9
        carrier = torch.empty(state.slab_width * xyz.shape[0], dtype=xyz.dtype, device=xyz.device)
        ctx.state = state
        ctx.has_sh = sh is not None
        ctx.save_for_backward(xyz, quaternion, scale, opacity_flat, sh, camera_T_world, K, centre)
        return uv, carrier

    @staticmethod
    def backward(ctx, grad_uv, grad_carrier):
        xyz, quaternion, scale, opacity_flat, sh, camera_T_world, K, centre = ctx.saved_tensors
        st = ctx.state
        N = st.N
        if grad_carrier is None:
            grad_carrier = torch.zeros(st.slab_width * N, dtype=xyz.dtype, device=xyz.device)
        slab = grad_carrier.contiguous()
        # Gradient on the projected means: the render backward's sums sit in the slab's uv section (by gaussian).
        # What autograd hands us for the COMPACT uv is either nothing / only what the caller added upstream of uv
        # (the render node did not emit its part: added to the slab's section in the kernel, row = rank among the
        # visible gaussians from the forward's scan), or the TOTAL (the render node did emit it because somebody
        # watches uv.grad: it then replaces the slab's section).  Incoming gradients are never edited.
        guv = None
        if grad_uv is not None and st.M > 0:
            guv = grad_uv.contiguous()
        with _stage(st, "preprocess_bwd"):
            grads = native().fused_preprocess_backward(slab, xyz, quaternion, scale, opacity_flat, sh,
                                                       camera_T_world, K, centre, st.visible, st.grad_out,
                                                       guv, st.scan if guv is not None else None,
                                                       not st.uv_grad_emitted)
        g_xyz, g_q, g_s, g_o, g_dc = grads[:5]
        g_sh = grads[5] if ctx.has_sh else None
        st.grad_flat = grads[-1]  # the one allocation all parameter gradients of this view are views of
        return g_xyz, g_q, g_s, g_o.view(-1, 1), g_dc, g_sh, None, None, None, None


class _CompositeTiles(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv, carrier, background_rgb, state):
        # the tile renderer was already enqueued by the projection node (right behind the binning, before the host
        # learned the pair count): this node only ties the image to uv and to the gradient carrier
        image, state.image = state.image, None
        ctx.state = state
        return image

    @staticmethod
    def backward(ctx, grad_image):
        st = ctx.state
        with _stage(st, "render_bwd"):
            slab = native().fused_render_backward(grad_image.contiguous(), st.N,
                                                  st.records if st.gather else st.stream_rec, st._ids_sorted,
                                                  st.ranges, st.background, st.n_per_pixel, st.w_per_pixel, st.masks,
                                                  st.keys_sorted, st.id_bits, st.gather)
        # The render pass's gradient on the compact uv is a gather of the slab's uv section (28 us at 3M).  It
        # reaches the per-gaussian backward through the slab anyway, so it is only materialised when it can be
        # OBSERVED: the caller retained uv's gradient (the reference trainer does, splat_py/trainer.py:360) or
        # hooked the tensor.
        uv_t = st.uv_ref() if st.uv_ref is not None else None
        watched = uv_t is None or uv_t.retains_grad or bool(uv_t._backward_hooks)
        st.uv_grad_emitted = bool(watched)
        if not watched:
            return None, slab, None, None
        N = st.N
        if st.slab_width == 12:   # interleaved rows: rgb3 opa | uv2 conic0 conic1 | conic2 pad3
            grad_uv = slab.view(N, 12).index_select(0, st.vis_idx)[:, 4:6].contiguous()
        else:                     # planar: rgb [N,3] | opacity [N] | uv [N,2] | conic [N,3]
            grad_uv = slab[4 * N:6 * N].view(N, 2).index_select(0, st.vis_idx)
        return grad_uv, slab, None, None


def rasterize(gaussians, camera_T_world, camera, near_thresh, far_thresh, cull_mask_padding, mh_dist,
              use_sh_precompute, background_rgb, return_state=False, profile=None, grad_out=None):
    if gaussians.sh is not None and not use_sh_precompute:
        return rasterize_unfused(gaussians, camera_T_world, camera, near_thresh, far_thresh,
                                 cull_mask_padding, mh_dist, use_sh_precompute, background_rgb)
    if gaussians.xyz.dtype != torch.float32:
        raise TypeError("the fused rasterizer is fp32; use rasterize_unfused for float64 inputs")
    state = _ViewState(profile)
    state.grad_out = grad_out
    state.background = background_rgb.contiguous()
    cfg = (int(camera.height), int(camera.width), float(near_thresh), float(far_thresh),
           float(cull_mask_padding), float(mh_dist))
    uv, carrier = _ProjectGaussians.apply(
        gaussians.xyz.contiguous(), gaussians.quaternion.contiguous(), gaussians.scale.contiguous(),
        gaussians.opacity.contiguous(), gaussians.rgb.contiguous(),
        None if gaussians.sh is None else gaussians.sh.contiguous(),
        camera_T_world.contiguous(), camera.K.contiguous(), state, cfg)
    state.uv_ref = weakref.ref(uv)
    image = _CompositeTiles.apply(uv, carrier, state.background, state)
    culling_mask = state.visible == 0
    if return_state:
        return image, culling_mask, uv, state
    return image, culling_mask, uv


def project_and_bin(gaussians, camera_T_world, camera, near_thresh, far_thresh, cull_mask_padding, mh_dist):
    """The per-gaussian part of the reference chain, one native op per reference op (any dtype):
    transform -> pinhole -> frustum cull -> Sigma_world / J / conic -> tile binning.  Returns a namespace with
    culling_mask [N] and, for the survivors, uv, xyz_cam, xyz, opacity (post-sigmoid), rgb, sh, conic,
    sorted_idx, tile_ranges."""
    from types import SimpleNamespace

    xyz_cam = transform_points_torch(gaussians.xyz, camera_T_world)
    uv = CameraPointProjection.apply(xyz_cam, camera.K)
    z, pad = xyz_cam[:, 2], cull_mask_padding
    culling_mask = (
        (z < near_thresh) | (z > far_thresh)
        | (uv[:, 0] < -1 * pad) | (uv[:, 0] > camera.width + pad)
        | (uv[:, 1] < -1 * pad) | (uv[:, 1] > camera.height + pad)
    )
    keep = ~culling_mask
    s = SimpleNamespace(culling_mask=culling_mask, uv=uv[keep, :], xyz_cam=xyz_cam[keep, :], xyz=gaussians.xyz[keep, :],
                        opacity=torch.sigmoid(gaussians.opacity[keep]), rgb=gaussians.rgb[keep, :],
                        sh=None if gaussians.sh is None else gaussians.sh[keep, :])
    sigma_world = ComputeSigmaWorld.apply(gaussians.quaternion[keep, :], gaussians.scale[keep, :])
    J = ComputeProjectionJacobian.apply(s.xyz_cam, camera.K)
    s.conic = ComputeConic.apply(sigma_world, J, camera_T_world)
    tiles = Tiles(camera.height, camera.width, uv.device)
    s.sorted_idx, s.tile_ranges = get_splats(s.uv.detach(), tiles, s.conic.detach(), s.xyz_cam.detach(), mh_dist)
    return s


def rasterize_unfused(gaussians, camera_T_world, camera, near_thresh, far_thresh, cull_mask_padding, mh_dist,
                      use_sh_precompute, background_rgb):
    """Operator-by-operator evaluation (any dtype, either SH mode); same return triple as rasterize()."""
    s = project_and_bin(gaussians, camera_T_world, camera, near_thresh, far_thresh, cull_mask_padding, mh_dist)
    rays = torch.zeros(1, 1, 1, dtype=gaussians.xyz.dtype, device=gaussians.xyz.device)  # unused unless per-pixel SH
    if s.sh is None:
        render_rgb = s.rgb
    else:
        coeffs = torch.cat((s.rgb.unsqueeze(dim=2), s.sh), dim=2)
        if use_sh_precompute:
            render_rgb = PrecomputeRGBFromSH.apply(coeffs, s.xyz, torch.inverse(camera_T_world).contiguous())
        else:
            render_rgb, rays = coeffs, compute_rays_in_world_frame(camera, camera_T_world)
    image = RenderImage.apply(render_rgb, s.opacity, s.uv, s.conic, rays, s.tile_ranges, s.sorted_idx,
                              (camera.height, camera.width), background_rgb)
    return image, s.culling_mask, s.uv

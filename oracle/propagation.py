"""CPU oracle for SURVEY row A10: the ProPainter flow-guided propagation prior.
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY UNPINNED.  The algorithm lives in an un-vendored, un-pinned third-party clone
(``calledit/DiffuEraser_np_array``, cloned at HEAD by
``/root/reference/install_videovanish.sh:78``; call site ``diffuerase.py:49-57``) that is
not under ``/root/reference``, and the reference holds no test or golden vector for it.
What follows restates the published ProPainter algorithm it wraps
(``propainter/model/propainter.py``: ``fbConsistencyCheck``, ``BidirectionalPropagation``
with ``learnable=False``, ``InpaintGenerator.img_propagation``;
``propainter/model/modules/flow_loss_utils.py``: ``flow_warp``;
``propainter/inference.py``: the sub-video loop with ``pad_len = 10``).

Two layers:

* ``*_torch`` functions run the restatement with torch's own CPU ``F.grid_sample`` so the
  gather / rounding semantics come from torch, not from us.
* ``model_propagate`` is an explicit float32 numpy model of the same scan (every rounding
  written out, FMA chain of torch's vectorised bilinear kernel emulated), operating on the
  packed u8 state the CUDA kernel uses.  ``tests/test_propagation_oracle.py`` checks it
  against the torch layer.
"""
import numpy as np

f32 = np.float32
HOLE = 1          # state bit0: pixel is still a hole
ZERO = 2          # state bit1: value is the float 0.0 of the masked frame, not a u8 level


# ======================================================================================
# torch restatement (upstream semantics)
# ======================================================================================

def flow_warp_torch(x, flow, interpolation="bilinear", padding_mode="zeros", align_corners=True):
    """flow_loss_utils.flow_warp: x (n,c,h,w), flow (n,h,w,2) in pixels."""
    import torch
    import torch.nn.functional as F
    _, _, h, w = x.size()
    grid_y, grid_x = torch.meshgrid(torch.arange(0, h), torch.arange(0, w), indexing="ij")
    grid = torch.stack((grid_x, grid_y), 2).type_as(x)
    grid_flow = grid + flow
    grid_flow_x = 2.0 * grid_flow[:, :, :, 0] / max(w - 1, 1) - 1.0
    grid_flow_y = 2.0 * grid_flow[:, :, :, 1] / max(h - 1, 1) - 1.0
    grid_flow = torch.stack((grid_flow_x, grid_flow_y), dim=3)
    return F.grid_sample(x, grid_flow, mode=interpolation, padding_mode=padding_mode, align_corners=align_corners)


def _length_sq(x):
    import torch
    return torch.sum(torch.square(x), dim=1, keepdim=True)


def fb_consistency_check_torch(flow_fw, flow_bw, alpha1=0.01, alpha2=0.5):
    """propainter.fbConsistencyCheck."""
    flow_bw_warped = flow_warp_torch(flow_bw, flow_fw.permute(0, 2, 3, 1))
    flow_diff_fw = flow_fw + flow_bw_warped
    mag_sq_fw = _length_sq(flow_fw) + _length_sq(flow_bw_warped)
    occ_thresh_fw = alpha1 * mag_sq_fw + alpha2
    return (_length_sq(flow_diff_fw) < occ_thresh_fw).to(flow_fw)


def _binary_mask(mask, th=0.1):
    mask[mask > th] = 1
    mask[mask <= th] = 0
    return mask


def bidirectional_propagation_torch(x, flows_forward, flows_backward, mask, interpolation="nearest"):
    """propainter.BidirectionalPropagation.forward with learnable=False.
    x [b,t,c,h,w], flows [b,t-1,2,h,w], mask [b,t,1,h,w] -> (outputs_b, outputs_f, masks_f)."""
    import torch
    b, t, c, h, w = x.shape
    feats = {"input": [x[:, i] for i in range(t)]}
    masks = {"input": [mask[:, i] for i in range(t)]}
    prop_list = ["backward_1", "forward_1"]
    cache_list = ["input"] + prop_list
    for p_i, module_name in enumerate(prop_list):
        feats[module_name], masks[module_name] = [], []
        if "backward" in module_name:
            frame_idx = list(range(t))[::-1]
            flow_idx = frame_idx
            flows_for_prop, flows_for_check = flows_forward, flows_backward
        else:
            frame_idx = list(range(t))
            flow_idx = list(range(-1, t - 1))
            flows_for_prop, flows_for_check = flows_backward, flows_forward
        for i, idx in enumerate(frame_idx):
            feat_current = feats[cache_list[p_i]][idx]
            mask_current = masks[cache_list[p_i]][idx]
            if i == 0:
                feat_prop, mask_prop = feat_current, mask_current
            else:
                flow_prop = flows_for_prop[:, flow_idx[i]]
                flow_check = flows_for_check[:, flow_idx[i]]
                flow_valid_mask = fb_consistency_check_torch(flow_prop, flow_check)
                feat_warped = flow_warp_torch(feat_prop, flow_prop.permute(0, 2, 3, 1), interpolation)
                mask_prop_valid = flow_warp_torch(mask_prop, flow_prop.permute(0, 2, 3, 1))
                mask_prop_valid = _binary_mask(mask_prop_valid)
                union_valid_mask = _binary_mask(mask_current * flow_valid_mask * (1 - mask_prop_valid))
                feat_prop = union_valid_mask * feat_warped + (1 - union_valid_mask) * feat_current
                mask_prop = _binary_mask(mask_current * (1 - (flow_valid_mask * (1 - mask_prop_valid))))
            feats[module_name].append(feat_prop)
            masks[module_name].append(mask_prop)
        if "backward" in module_name:
            feats[module_name] = feats[module_name][::-1]
            masks[module_name] = masks[module_name][::-1]
    outputs_b = torch.stack(feats["backward_1"], dim=1)
    outputs_f = torch.stack(feats["forward_1"], dim=1)
    masks_f = torch.stack(masks["forward_1"], dim=1)
    return outputs_b, outputs_f, masks_f


def to_unit_float(frames_u8):
    """to_tensors() then ``* 2 - 1`` (inference.py): u8 -> f32 in [-1, 1]."""
    return ((frames_u8.astype(f32) / f32(255.0)).astype(f32) * f32(2.0) - f32(1.0)).astype(f32)


def img_propagation_torch(frames_u8, masks_u8, flows_f, flows_b):
    """One sub-video through ``img_propagation(..., 'nearest')`` + the compose line
    ``updated = frames*(1-m) + prop*m``.  frames u8 [T,h,w,3], masks u8 [T,h,w] (>0 = hole),
    flows f32 [T-1,h,w,2].  Returns (updated_frames f32 [T,3,h,w], updated_masks f32 [T,h,w])."""
    import torch
    frames = torch.from_numpy(to_unit_float(frames_u8)).permute(0, 3, 1, 2)[None]      # [1,T,3,h,w]
    m = torch.from_numpy((masks_u8 > 0).astype(f32))[None, :, None]                     # [1,T,1,h,w]
    ff = torch.from_numpy(flows_f).permute(0, 3, 1, 2)[None]
    fb = torch.from_numpy(flows_b).permute(0, 3, 1, 2)[None]
    masked = frames * (1 - m)
    _, prop, upd_m = bidirectional_propagation_torch(masked, ff, fb, m.clone(), "nearest")
    updated = frames * (1 - m) + prop * m
    return updated[0].numpy(), upd_m[0, :, 0].numpy()


def subvideo_plan(video_length, subvideo_length=50, pad_len=10):
    """inference.py sub-video loop: list of (s_f, e_f, pad_len_s, pad_len_e).  The caller
    keeps frames [pad_len_s, e_f - s_f - pad_len_e) of each window.  diffuerase.py:54 passes
    subvideo_length=50; the image-propagation window is min(100, subvideo_length)."""
    sub = min(100, subvideo_length)
    if video_length <= sub:
        return [(0, video_length, 0, 0)]
    plan = []
    for f in range(0, video_length, sub):
        s_f = max(0, f - pad_len)
        e_f = min(video_length, f + sub + pad_len)
        plan.append((s_f, e_f, max(0, f) - s_f, e_f - min(video_length, f + sub)))
    return plan


def propagate_clip_torch(frames_u8, masks_u8, flows_f, flows_b, subvideo_length=50, pad_len=10):
    """Whole clip: sub-video windows, pads discarded, results concatenated."""
    outs, ms = [], []
    for s_f, e_f, ps, pe in subvideo_plan(len(frames_u8), subvideo_length, pad_len):
        u, m = img_propagation_torch(frames_u8[s_f:e_f], masks_u8[s_f:e_f], flows_f[s_f:e_f - 1], flows_b[s_f:e_f - 1])
        outs.append(u[ps:e_f - s_f - pe])
        ms.append(m[ps:e_f - s_f - pe])
    return np.concatenate(outs), np.concatenate(ms)


# ======================================================================================
# explicit numpy model on the packed u8 state (the kernel's arithmetic spec)
# ======================================================================================

def pack_state(frames_u8, masks_u8):
    """u8 [T,h,w,3] + mask -> u32 [T,h,w]: R | G<<8 | B<<16 | state<<24.  Holes carry the
    masked frame's 0.0: rgb = 0, state = HOLE|ZERO."""
    hole = masks_u8 > 0
    p = (frames_u8[..., 0].astype(np.uint32) | (frames_u8[..., 1].astype(np.uint32) << 8) |
         (frames_u8[..., 2].astype(np.uint32) << 16))
    return np.where(hole, np.uint32((HOLE | ZERO) << 24), p).astype(np.uint32)


def decode_state(packed):
    """packed u32 [T,h,w] -> (frames f32 [T,3,h,w] in [-1,1] with ZERO pixels = 0.0,
    hole mask f32 [T,h,w])."""
    rgb = np.stack([(packed >> s) & 0xFF for s in (0, 8, 16)], axis=1).astype(np.uint8)
    val = to_unit_float(rgb)
    zero = ((packed >> 24) & ZERO) != 0
    val = np.where(zero[:, None], f32(0), val).astype(f32)
    return val, (((packed >> 24) & HOLE) != 0).astype(f32)


def _fma(a, b, c):
    """float32 fused multiply-add: the product of two f32 is exact in f64."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def _sample_coords(flow, h, w):
    """Un-normalised sample position of flow_warp + grid_sample(align_corners=True)."""
    gy, gx = np.mgrid[0:h, 0:w]
    px = (gx.astype(f32) + flow[..., 0]).astype(f32)
    py = (gy.astype(f32) + flow[..., 1]).astype(f32)
    nx = ((f32(2.0) * px).astype(f32) / f32(max(w - 1, 1)) - f32(1.0)).astype(f32)
    ny = ((f32(2.0) * py).astype(f32) / f32(max(h - 1, 1)) - f32(1.0)).astype(f32)
    ix = ((nx + f32(1.0)).astype(f32) * f32((w - 1) / 2.0)).astype(f32)
    iy = ((ny + f32(1.0)).astype(f32) * f32((h - 1) / 2.0)).astype(f32)
    return ix, iy


def _bilinear(chans, ix, iy):
    """torch CPU vectorised bilinear, zeros padding: r = a*nw; r = fma(b,ne,r); fma(c,sw,r); fma(d,se,r)."""
    h, w = ix.shape
    x0f, y0f = np.floor(ix), np.floor(iy)
    wx, wy = (ix - x0f).astype(f32), (iy - y0f).astype(f32)
    ex, sy = (f32(1) - wx).astype(f32), (f32(1) - wy).astype(f32)
    nw, ne, sw, se = (sy * ex).astype(f32), (sy * wx).astype(f32), (wy * ex).astype(f32), (wy * wx).astype(f32)
    big = 1 << 30
    x0 = np.clip(x0f, -big, big).astype(np.int64)
    y0 = np.clip(y0f, -big, big).astype(np.int64)

    def tap(img, yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        return np.where(ok, img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], f32(0)).astype(f32)

    outs = []
    for img in chans:
        r = (tap(img, y0, x0) * nw).astype(f32)
        r = _fma(tap(img, y0, x0 + 1), ne, r)
        r = _fma(tap(img, y0 + 1, x0), sw, r)
        r = _fma(tap(img, y0 + 1, x0 + 1), se, r)
        outs.append(r)
    return outs


def _step(cur, prev, flow_prop, flow_check):
    """One propagation step on packed states (u32 [h,w])."""
    h, w = cur.shape
    ix, iy = _sample_coords(flow_prop, h, w)
    bwx, bwy = _bilinear([flow_check[..., 0], flow_check[..., 1]], ix, iy)
    fx, fy = flow_prop[..., 0], flow_prop[..., 1]
    dx, dy = (fx + bwx).astype(f32), (fy + bwy).astype(f32)
    diff = ((dx * dx).astype(f32) + (dy * dy).astype(f32)).astype(f32)
    mag = (((fx * fx).astype(f32) + (fy * fy).astype(f32)).astype(f32) +
           ((bwx * bwx).astype(f32) + (bwy * bwy).astype(f32)).astype(f32)).astype(f32)
    thr = ((f32(0.01) * mag).astype(f32) + f32(0.5)).astype(f32)
    valid = diff < thr
    prev_hole = (((prev >> 24) & HOLE) != 0).astype(f32)
    (mpv,) = _bilinear([prev_hole], ix, iy)
    mpv = mpv > f32(0.1)
    cur_hole = ((cur >> 24) & HOLE) != 0
    fill = cur_hole & valid & ~mpv
    big = 1 << 30
    xi = np.clip(np.rint(ix), -big, big).astype(np.int64)
    yi = np.clip(np.rint(iy), -big, big).astype(np.int64)
    inb = (yi >= 0) & (yi < h) & (xi >= 0) & (xi < w)
    src = prev[np.clip(yi, 0, h - 1), np.clip(xi, 0, w - 1)]
    warped = np.where(inb, src & np.uint32(~(HOLE << 24) & 0xFFFFFFFF), np.uint32(ZERO << 24)).astype(np.uint32)
    # a hole source cannot be selected (its bilinear weight is >= 0.25 > 0.1); keep its value anyway
    return np.where(fill, warped, cur).astype(np.uint32)


def model_propagate(frames_u8, masks_u8, flows_f, flows_b):
    """Explicit model of img_propagation for one sub-video -> packed u32 [T,h,w] (forward pass)."""
    t = len(frames_u8)
    inp = pack_state(frames_u8, masks_u8)
    back = [None] * t
    for i, idx in enumerate(range(t - 1, -1, -1)):
        back[idx] = inp[idx] if i == 0 else _step(inp[idx], back[idx + 1], flows_f[idx], flows_b[idx])
    fwd = [None] * t
    for i in range(t):
        fwd[i] = back[i] if i == 0 else _step(back[i], fwd[i - 1], flows_b[i - 1], flows_f[i - 1])
    return np.stack(fwd)


def model_propagate_clip(frames_u8, masks_u8, flows_f, flows_b, subvideo_length=50, pad_len=10):
    outs = []
    for s_f, e_f, ps, pe in subvideo_plan(len(frames_u8), subvideo_length, pad_len):
        p = model_propagate(frames_u8[s_f:e_f], masks_u8[s_f:e_f], flows_f[s_f:e_f - 1], flows_b[s_f:e_f - 1])
        outs.append(p[ps:e_f - s_f - pe])
    return np.concatenate(outs)


# ======================================================================================
# N2: the pixel glue after the ProPainter network (neighbour windows, 0.5 / 0.5 merge)
# ======================================================================================

def get_ref_index(mid_neighbor_id, neighbor_ids, length, ref_stride=10, ref_num=-1):
    """propainter/inference.py get_ref_index: the non-local reference frames of one window."""
    ref_index = []
    if ref_num == -1:
        for i in range(0, length, ref_stride):
            if i not in neighbor_ids:
                ref_index.append(i)
    else:
        start_idx = max(0, mid_neighbor_id - ref_stride * (ref_num // 2))
        end_idx = min(length, mid_neighbor_id + ref_stride * (ref_num // 2))
        for i in range(start_idx, end_idx, ref_stride):
            if i not in neighbor_ids:
                if len(ref_index) > ref_num:
                    break
                ref_index.append(i)
    return ref_index


def neighbor_plan(video_length, neighbor_length=10, ref_stride=10, subvideo_length=50):
    """The sliding windows of the feature-propagation + transformer loop of propainter/inference.py:
    [(neighbor_ids, ref_ids)] for f in range(0, video_length, neighbor_length // 2)
    (diffuerase.py:54 passes ref_stride=10, neighbor_length=10, subvideo_length=50)."""
    neighbor_stride = neighbor_length // 2
    ref_num = subvideo_length // ref_stride if video_length > subvideo_length else -1
    plan = []
    for f in range(0, video_length, neighbor_stride):
        ids = list(range(max(0, f - neighbor_stride), min(video_length, f + neighbor_stride + 1)))
        plan.append((ids, get_ref_index(f, ids, video_length, ref_stride, ref_num)))
    return plan


def ref_neighbor_merge(pred_windows, plan, masks01, ori_frames):
    """The compose loop of propainter/inference.py restated with numpy.  pred_windows[k] = network output
    f32 [len(neighbor_ids_k), 3, h, w] in [-1, 1]; masks01 u8 [T,h,w] in {0,1}; ori_frames u8 [T,h,w,3].
    Returns the list of composed u8 frames."""
    comp = [None] * len(ori_frames)
    for (ids, _), pred in zip(plan, pred_windows):
        pred_img = ((pred.astype(f32) + f32(1)) / f32(2)).astype(f32)
        pred_img = np.transpose(pred_img, (0, 2, 3, 1)) * 255          # float32 * python int stays float32
        for i, idx in enumerate(ids):
            m = masks01[idx][..., None]
            img = np.array(pred_img[i]).astype(np.uint8) * m + ori_frames[idx] * (1 - m)
            if comp[idx] is None:
                comp[idx] = img
            else:
                comp[idx] = comp[idx].astype(np.float32) * 0.5 + img.astype(np.float32) * 0.5
            comp[idx] = comp[idx].astype(np.uint8)
    return comp

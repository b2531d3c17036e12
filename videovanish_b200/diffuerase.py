"""Drop-in for the reference's ``diffuerase`` module (/root/reference/diffuerase.py).

``run_infill_on_frames`` keeps the reference's signature, progress milestones, model hand-off
contract and return convention (the list returned by the model's ``forward``,
post-processed in place, C-contiguous uint8 HxWx3 arrays).  The two pixel loops
(diffuerase.py:28-31 and :70-112) run on the GPU through ``libvvb200.so``; the DiffuEraser /
ProPainter networks are out of scope and are used exactly as the reference uses them (lazy import
of the un-vendored packages, module-level cache), or injected with ``set_models`` for tests.

Differences, all deliberate and documented in DESIGN.md:
* diffuerase.py:114 returns inside the ``for`` loop, so the reference post-processes frame 0 only.
  Here every frame is processed (the evident intent); ``BUG_COMPAT = True`` restores the literal
  behaviour.
* There is no CPU fallback: without a CUDA device the call raises.

When every model in play offers ``forward_device`` (the adapters of ``wrappers.py`` around the networks),
the call is device-resident: frames and masks are uploaded once, K1 -> [prior stages] -> [DiffuEraser
wrapper stages] -> K3 run back to back in HBM, and only the finished frames come back.
"""
import math

import numpy as np

from . import hostpipe

BUG_COMPAT = False           # True: reproduce the early return of diffuerase.py:114
# Device-resident route: when the dilated masks leave most rows untouched, only the rows they reach are downloaded
# and the rest of each finished frame is copied from the caller's input frame on the host (same bytes, see
# _run_on_device).  Used when the rows to download are at most this fraction of the clip.
ROW_BOUNDED_RESULTS = True
ROW_BOUNDED_MAX_FRACTION = 0.75
last_call_info = {}          # what the last device-resident call did: {"row_bounded": bool, "rows_downloaded": int}

device = None
last_ckpt = None
video_inpainting_sd = None
propainter = None
_pipeline = None


def set_models(diffueraser=None, propainter_model=None):
    """Inject model objects exposing the upstream ``forward`` signatures (tests, benchmarks)."""
    global video_inpainting_sd, propainter, last_ckpt
    if diffueraser is not None:
        video_inpainting_sd = diffueraser
        last_ckpt = "2-Step"
    if propainter_model is not None:
        propainter = propainter_model


def _get_pipeline(h0, w0):
    global _pipeline
    if _pipeline is None or _pipeline.geometry != (h0, w0):
        if _pipeline is not None:
            _pipeline.close()
        _pipeline = hostpipe.HostPipeline(h0, w0)
    return _pipeline


def run_infill_on_frames(frames_rgb, mask_frames, mask_dilation_iter=8, ckpt="2-Step",
                         propainer_frames=None, max_img_size=960, keep_unmasked_original=True, feather_px=3,
                         prog=None):
    """Same contract as reference diffuerase.py:20-114."""
    global device, last_ckpt, video_inpainting_sd, propainter

    H0, W0 = _frame_size(frames_rgb)
    pipe = _get_pipeline(H0, W0)

    if _device_route(propainer_frames):
        return _run_on_device(pipe, frames_rgb, mask_frames, mask_dilation_iter, propainer_frames, max_img_size,
                              keep_unmasked_original, feather_px, prog)

    if prog is not None: prog(5, "dilating frames")
    mask_size = _frame_size(mask_frames)
    if mask_size == (H0, W0):
        dilated_mask_frames = pipe.pre(mask_frames, mask_dilation_iter)                  # :27-31 (K1)
    else:
        # masks of another size than the frames: K1 at the masks' own size (what the models get, like in the reference);
        # the post stage fits them to the frames with INTER_NEAREST (:85-86)
        mask_pipe = hostpipe.HostPipeline(*mask_size)
        try:
            dilated_mask_frames = mask_pipe.pre(mask_frames, mask_dilation_iter)
        finally:
            mask_pipe.close()

    if prog is not None: prog(10, "loading weights")
    if last_ckpt != ckpt:                                                                # :35-45
        from diffueraser.diffueraser import DiffuEraser
        from propainter.inference import get_device
        device = get_device()
        ckpt = "2-Step"
        last_ckpt = ckpt
        video_inpainting_sd = DiffuEraser(device, "stable-diffusion-v1-5/stable-diffusion-v1-5",
                                          "stabilityai/sd-vae-ft-mse", "lixiaowen/diffuEraser", ckpt=ckpt)

    if propainer_frames is None:                                                         # :47-57
        if propainter is None:
            from propainter.inference import Propainter
            propainter = Propainter("ruffy369/propainter", device=device)
        if prog is not None: prog(20, "running propainter prior")
        propainer_frames = propainter.forward(frames_rgb, dilated_mask_frames, ref_stride=10, neighbor_length=10,
                                              subvideo_length=50, mask_dilation=0, progress=prog)

    if prog is not None: prog(50, "running DiffuEraser")
    guidance_scale = None
    inpainted_frames = video_inpainting_sd.forward(frames_rgb, dilated_mask_frames, propainer_frames,   # :62-67
                                                   max_img_size=max_img_size, mask_dilation_iter=0,
                                                   guidance_scale=guidance_scale, progress=prog)

    if prog is not None: prog(90, "resizing and merging finished frames")
    n = 1 if BUG_COMPAT else len(inpainted_frames)                                       # :114
    if n == 0:
        return inpainted_frames
    fh, fw = inpainted_frames[0].shape[:2]
    if (fh, fw) == (H0, W0) and not keep_unmasked_original:
        return inpainted_frames                                                          # nothing to do (:72, :75)
    if fh * fw > H0 * W0:             # never produced by the model wrapper (row A9 only shrinks)
        raise ValueError("inpainted frames (%dx%d) larger than the originals (%dx%d)" % (fh, fw, H0, W0))
    # the dilated masks are still on the device from `pre` (unless they had to be fitted to the frames' size)
    fitted = None
    if mask_size != (H0, W0) and keep_unmasked_original:
        fitted = _fit_masks(dilated_mask_frames[:n], H0, W0)                                          # :85-86 (K2 NEAREST)
    out = pipe.post(inpainted_frames[:n], frames_rgb[:n], fitted, feather_px, keep_unmasked_original)   # :70-112 (K3)
    for i in range(n):
        inpainted_frames[i] = out[i]
    return inpainted_frames


def _fit_masks(dilated, h0, w0):
    """diffuerase.py:85-86: dilated masks of another size -> the frames' size with cv2.INTER_NEAREST semantics (K2)."""
    import torch

    from . import ops
    d = torch.from_numpy(np.ascontiguousarray(np.stack(dilated))).cuda()
    return list(ops.resize(d, h0, w0, ops.INTER_NEAREST).cpu().numpy())


def _frame_size(frames):
    """(H0, W0) of a list of host frames or of a tools.DeviceFrames clip (without downloading it)."""
    return tuple(frames.tensor.shape[1:3]) if hasattr(frames, "tensor") else tuple(frames[0].shape[:2])


def _frame_shape(frames):
    return tuple(frames.tensor.shape[1:]) if hasattr(frames, "tensor") else tuple(frames[0].shape)


def _device_route(propainer_frames):
    """True when the models that this call will use are the device-resident adapters."""
    if video_inpainting_sd is None or last_ckpt != "2-Step" or not hasattr(video_inpainting_sd, "forward_device"):
        return False
    return propainer_frames is not None or (propainter is not None and hasattr(propainter, "forward_device"))


def _run_on_device(pipe, frames_rgb, mask_frames, mask_dilation_iter, propainer_frames, max_img_size,
                   keep_unmasked_original, feather_px, prog, upload_pipe=None, keep_on_device=False):
    """Same stages, milestones and results as the host-list route of ``run_infill_on_frames``, with the clip
    resident in HBM from the first upload to the last download.  ``upload_pipe``: a second pipeline for the
    uploads, so that they do not queue behind another chunk's download; ``keep_on_device``: return the u8
    [T,H0,W0,3] device tensor instead of downloading it (the chunked driver cross-fades overlaps in HBM)."""
    import torch

    from . import ops, wrappers
    H0, W0 = _frame_size(frames_rgb)
    t = len(frames_rgb)
    resident = hasattr(frames_rgb, "tensor")            # tools.DeviceFrames in -> tools.DeviceFrames out
    up = upload_pipe if upload_pipe is not None else pipe
    with torch.cuda.device(pipe.device):
        if prog is not None: prog(5, "dilating frames")
        masks = up.upload(mask_frames, _frame_shape(mask_frames), is_mask=True)
        h, w = ops.inference_size(H0, W0, max_img_size)
        if tuple(masks.shape[1:3]) == (H0, W0):
            dil, low, bits = ops.binarize_dilate(masks, mask_dilation_iter, lowres_size=(h, w), return_bits=True)   # K1
        else:
            # masks of another size than the frames: K1 at their own size; the model-side mask is their INTER_NEAREST
            # down-size (straight from that size), the post stage gets them fitted to the frames' size (:85-86)
            dil, low = ops.binarize_dilate(masks, mask_dilation_iter, lowres_size=(h, w))
            dil = ops.resize(dil, H0, W0, ops.INTER_NEAREST)
            bits = None
        del masks
        # Row-bounded result: outside the rows a frame's dilated mask (+ feather radius) reaches, the finished frame IS
        # the input frame (alpha = 0 -> rint(orig) = orig, diffuerase.py:112), so those rows are copied host -> host from
        # the caller's frames in the background and only the rows in between come back over PCIe.
        rows = None
        if (ROW_BOUNDED_RESULTS and keep_unmasked_original and not BUG_COMPAT and not resident and not keep_on_device
                and bits is not None and all(isinstance(f, np.ndarray) and f.dtype == np.uint8 and f.flags.c_contiguous
                                             and f.shape == (H0, W0, 3) for f in frames_rgb)):
            margin = max(0, int(math.ceil(float(feather_px))) - 1) + 1 if feather_px > 0 else 1
            bounds = ops.mask_row_bounds(bits, margin).cpu().numpy()          # tiny; waits for the mask upload + K1 only
            lo, hi = bounds[:, 0], bounds[:, 1]
            if int((hi - lo).sum()) <= ROW_BOUNDED_MAX_FRACTION * t * H0:
                from .hostpipe import pinned_frames
                result = pinned_frames(t, (H0, W0, 3))
                pipe.host_rows_begin(result, list(frames_rgb), lo, hi)
                rows = (result, lo, hi)
        last_call_info["row_bounded"] = rows is not None
        last_call_info["rows_downloaded"] = int((rows[2] - rows[1]).sum()) if rows is not None else t * H0
        frames = up.upload(frames_rgb, (H0, W0, 3))
        clip = wrappers.DeviceClip(frames, dil, lowres=low)
        clip.mask_bits = bits
        if prog is not None: prog(10, "loading weights")
        if propainer_frames is None:
            if prog is not None: prog(20, "running propainter prior")
            priors = propainter.forward_device(clip, ref_stride=10, neighbor_length=10, subvideo_length=50,
                                               mask_dilation=0, progress=prog)
        else:
            priors = up.upload(propainer_frames, _frame_shape(propainer_frames))
        if prog is not None: prog(50, "running DiffuEraser")
        inpainted = video_inpainting_sd.forward_device(clip, priors, max_img_size=max_img_size, mask_dilation_iter=0,
                                                       guidance_scale=None, progress=prog)
        del priors
        if prog is not None: prog(90, "resizing and merging finished frames")
        n = min(1, t) if BUG_COMPAT else t
        fh, fw = inpainted.shape[1:3]
        if (fh, fw) == (H0, W0) and not keep_unmasked_original:
            out = inpainted
        else:
            out = ops.upscale_feather_composite(inpainted[:n], frames[:n], dil[:n], feather_px, keep_unmasked_original,
                                                mask_bits=None if bits is None else bits[:n])        # K3
        if keep_on_device and n == t:
            return out
        if resident and n == t:
            from .tools import DeviceFrames
            return DeviceFrames(out)
        if rows is not None and tuple(out.shape) == (t, H0, W0, 3):
            return pipe.download_rows(out, *rows)
        result = pipe.download(out)
        if n < t and out is not inpainted:
            result += pipe.download(inpainted[n:])
        return result


_upload_pipeline = None


def run_infill_on_frames_chunked(frames_rgb, mask_frames, chunk=80, overlap=16, **kwargs):
    """Long clips in overlapping chunks (the feature the reference advertises at README.md:18 and lists as
    a TODO at README.md:76; builder-defined spec): every chunk of ``chunk`` frames goes
    through ``run_infill_on_frames`` on its own (so the models only ever see ``chunk`` frames), and the
    ``overlap`` frames shared by consecutive chunks are cross-faded on the GPU by K5
    (``w = (k+1)/(overlap+1)``).  ``propainer_frames``, if given, is sliced per chunk.

    With the device-resident adapters installed the chunks never leave HBM between their stages and the
    cross-fade: chunk c+1 is uploaded and computed while chunk c's finished frames are downloaded (two pipelines,
    PCIe both ways at once), and the shared frames are blended in place on the device."""
    import torch

    from . import chunking, hostpipe, ops
    global _upload_pipeline
    n = len(frames_rgb)
    plan = chunking.chunk_plan(n, chunk, overlap)
    priors = kwargs.pop("propainer_frames", None)
    result = [None] * n

    if _device_route(priors) and not BUG_COMPAT:
        H0, W0 = _frame_size(frames_rgb)
        pipe = _get_pipeline(H0, W0)
        if _upload_pipeline is None or _upload_pipeline.geometry != (H0, W0):
            _upload_pipeline = hostpipe.HostPipeline(H0, W0)
        args = dict(mask_dilation_iter=kwargs.get("mask_dilation_iter", 8), max_img_size=kwargs.get("max_img_size", 960),
                    keep_unmasked_original=kwargs.get("keep_unmasked_original", True), feather_px=kwargs.get("feather_px", 3),
                    prog=kwargs.get("prog"))
        dl_stream = torch.cuda.Stream()
        pending = None                       # (device frames to emit, first clip index, event) of the previous chunk

        def emit(item):
            tensor, first, ev = item
            with torch.cuda.stream(dl_stream):
                dl_stream.wait_event(ev)
                frames = pipe.download(tensor)
            result[first:first + len(frames)] = frames

        prev_tail = None
        for ci, (s, e) in enumerate(plan):
            out = _run_on_device(pipe, frames_rgb[s:e], mask_frames[s:e], args["mask_dilation_iter"],
                                 None if priors is None else priors[s:e], args["max_img_size"],
                                 args["keep_unmasked_original"], args["feather_px"], args["prog"],
                                 upload_pipe=_upload_pipeline, keep_on_device=True)
            if ci > 0:
                ov = plan[ci - 1][1] - s
                ops.chunk_blend(prev_tail[prev_tail.shape[0] - ov:], out[:ov], out=out[:ov])      # in place, in HBM
            nxt_ov = (e - plan[ci + 1][0]) if ci + 1 < len(plan) else 0
            keep = (e - s) - nxt_ov
            prev_tail = out[keep:] if nxt_ov else None
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                emit(pending)                # blocks on the previous chunk's download while this chunk uploads / computes
            pending = (out[:keep], s, ev)
        emit(pending)
        return result

    prev_tail = None                     # device copy of the previous chunk's last `overlap` frames
    pipe = None
    for ci, (s, e) in enumerate(plan):
        out = run_infill_on_frames(frames_rgb[s:e], mask_frames[s:e],
                                   propainer_frames=None if priors is None else priors[s:e], **kwargs)
        if pipe is None:
            pipe = _get_pipeline(*out[0].shape[:2])
        shape = tuple(out[0].shape)
        lo = 0
        if ci > 0:
            ov = plan[ci - 1][1] - s
            head = pipe.upload(out[:ov], shape)
            blended = pipe.download(ops.chunk_blend(prev_tail[prev_tail.shape[0] - ov:], head))
            for k in range(ov):
                result[s + k] = blended[k]
            lo = ov
        for k in range(lo, e - s):
            result[s + k] = out[k]
        if ci + 1 < len(plan):
            nxt_ov = e - plan[ci + 1][0]
            prev_tail = pipe.upload(out[(e - s) - nxt_ov:], shape)
    return result


def main():
    """CLI of reference diffuerase.py:121-151 (same flags).  The reference's inverted
    ``--prior_video`` test (:142) is NOT reproduced: the prior is loaded when it is given."""
    import argparse
    import os

    from . import tools
    ap = argparse.ArgumentParser(description="Vanish masked objects from a video (B200 pixel pipeline).")
    ap.add_argument("--color_video", required=True, type=str)
    ap.add_argument("--mask_video", required=True, type=str)
    ap.add_argument("--prior_video", required=False, type=str)
    ap.add_argument("--start_frame", type=int, default=0)
    ap.add_argument("--max_frames", type=int, default=-1)
    ap.add_argument("--out", type=str, default=None)
    args = ap.parse_args()
    assert os.path.isfile(args.color_video), "input video missing"
    out_video = args.out or (args.color_video + "_vanished.mkv")
    frames, fps = tools.load_video_frames_from_path(args.color_video, args.start_frame, args.max_frames)
    H0, W0 = frames[0].shape[:2]
    mask_frames, _ = tools.load_video_frames_from_path(args.mask_video, args.start_frame, args.max_frames)
    assert mask_frames[0].shape[:2] == (H0, W0), "mask and color video are diffrent sizes"
    prior_frames = None
    if args.prior_video is not None:
        prior_frames, _ = tools.load_video_frames_from_path(args.prior_video, args.start_frame, args.max_frames)
        assert prior_frames[0].shape[:2] == (H0, W0), "prior and color video are diffrent sizes"
    result = run_infill_on_frames(frames, mask_frames, propainer_frames=prior_frames)
    tools.write_video_frames_to_path(out_video, result, fps, H0, W0)


if __name__ == "__main__":
    main()

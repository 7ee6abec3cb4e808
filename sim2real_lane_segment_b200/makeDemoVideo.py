"""GPU mirror of the reference's `makeDemoVideo.py` (predictVideo, lines 15-47) and of the overlay in `comparison.py`.

`predictFrames(frames, model, transform)` is the per-frame body of the reference loop for a whole batch of frames kept on
the GPU: MyTransform -> model forward -> argmax mask -> frame resized to the network resolution with the classes painted
over it (`rl_overlay`).  `predictVideo` keeps the reference's signature and does the container I/O with cv2 on the host."""
import ctypes

import numpy as np
import torch

from . import _lib


def predictFrames(frames, model, transform):
    """frames: uint8 [B,Hs,Ws,3] BGR (numpy or tensor).  Returns (overlay uint8 [B,h,w,3] on the GPU, mask uint8 [B,h,w])"""
    L = _lib.lib()
    fr = torch.as_tensor(np.ascontiguousarray(frames) if isinstance(frames, np.ndarray) else frames)
    if fr.dim() == 3:
        fr = fr[None]
    fr = fr.cuda(non_blocking=True).contiguous()
    x, _ = transform(fr)
    was_training = model.training
    model.eval()
    with torch.no_grad():
        tok = model.featureExtractor(x)
        _, mask = model.runtime().run_head(tok, want_probs=False, want_mask=True)
    model.train(was_training)
    B, Hs, Ws, _ = fr.shape
    h, w = x.shape[-2:]
    out = torch.empty(B, h, w, 3, dtype=torch.uint8, device="cuda")
    rc = L.rl_overlay(ctypes.c_void_p(fr.data_ptr()), B, Hs, Ws, h, w, ctypes.c_void_p(mask.data_ptr()), 0,
                      ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        L.rl_last_error.restype = ctypes.c_char_p
        raise RuntimeError(L.rl_last_error().decode())
    return out, mask


def predictVideo(inputVideo, outputVideo, model, transform, batch=32):
    import cv2
    cap_in = cv2.VideoCapture(inputVideo)
    fourcc = cv2.VideoWriter_fourcc(*'FFV1')
    fps = cap_in.get(cv2.CAP_PROP_FPS)
    w_out = cv2.VideoWriter(outputVideo, fourcc, fps, (transform.width, transform.height), True)
    try:
        pending = []

        def flush():
            if pending:
                out, _ = predictFrames(np.stack(pending), model, transform)
                for f in out.cpu().numpy():
                    w_out.write(f)
                pending.clear()

        for _ in range(int(cap_in.get(cv2.CAP_PROP_FRAME_COUNT))):
            ret, frame = cap_in.read()
            if not ret:
                continue
            pending.append(frame)
            if len(pending) == batch:
                flush()
        flush()
    finally:
        cap_in.release()
        w_out.release()

"""Host -> device staging for callers that hand pageable CPU tensors to the loss / renderer.

``tensor.cuda()`` on pageable memory is a blocking copy through the driver's own small bounce buffer.  Here the
tensor is cut into chunks that are copied by the CPU (torch's multi-threaded ``copy_``) into one of two pinned
buffers and sent with ``cudaMemcpyAsync`` on the current stream, so the CPU copy of chunk ``i+1`` overlaps the DMA
of chunk ``i`` and the transfer runs at min(host memcpy, PCIe) instead of their sum.  Stream-ordered: kernels
enqueued afterwards on the current stream see the data; the host returns when the last chunk has been *enqueued*.
"""
import threading

import torch

CHUNK_BYTES = 32 << 20
_lock = threading.Lock()
_state = {}     # device index -> (pinned buffers, events)


def _buffers(index):
    st = _state.get(index)
    if st is None:
        bufs = [torch.empty(CHUNK_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        evs = [torch.cuda.Event() for _ in range(2)]
        st = _state[index] = (bufs, evs)
    return st


def upload(t, device=None):
    """Contiguous CPU tensor -> new CUDA tensor with the same shape and dtype on ``device`` (default: current)."""
    if t.is_cuda:
        return t
    if not t.is_contiguous():
        t = t.contiguous()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty(t.shape, dtype=t.dtype, device=device)
    nbytes = t.numel() * t.element_size()
    if nbytes == 0:
        return out
    src = t.detach().view(-1).view(torch.uint8)
    dst = out.view(-1).view(torch.uint8)
    with _lock, torch.cuda.device(device):
        bufs, evs = _buffers(device.index)
        for i, lo in enumerate(range(0, nbytes, CHUNK_BYTES)):
            n = min(CHUNK_BYTES, nbytes - lo)
            k = i & 1
            evs[k].synchronize()                               # the DMA that last read this pinned buffer is done
            bufs[k][:n].copy_(src[lo:lo + n])                  # CPU copy into pinned memory (overlaps the previous DMA)
            dst[lo:lo + n].copy_(bufs[k][:n], non_blocking=True)
            evs[k].record()
    return out


def download(t):
    """CUDA tensor -> new (pageable) CPU tensor, the mirror of :func:`upload`: chunk ``i`` is DMA'd into one pinned buffer
    while the CPU copies chunk ``i-1`` out of the other.  Blocks until the data is on the host."""
    if not t.is_cuda:
        return t
    t = t.detach().contiguous()
    out = torch.empty(t.shape, dtype=t.dtype)
    nbytes = t.numel() * t.element_size()
    if nbytes == 0:
        return out
    src = t.view(-1).view(torch.uint8)
    dst = out.view(-1).view(torch.uint8)
    with _lock, torch.cuda.device(t.device):
        bufs, evs = _buffers(t.device.index)
        for ev in evs:
            ev.synchronize()
        chunks = [(lo, min(CHUNK_BYTES, nbytes - lo)) for lo in range(0, nbytes, CHUNK_BYTES)]
        for i, (lo, n) in enumerate(chunks):
            k = i & 1
            bufs[k][:n].copy_(src[lo:lo + n], non_blocking=True)
            evs[k].record()
            if i > 0:
                plo, pn = chunks[i - 1]
                evs[1 - k].synchronize()
                dst[plo:plo + pn].copy_(bufs[1 - k][:pn])
        plo, pn = chunks[-1]
        k = (len(chunks) - 1) & 1
        evs[k].synchronize()
        dst[plo:plo + pn].copy_(bufs[k][:pn])
    return out


class _StagedUpload(torch.autograd.Function):
    """Differentiable ``upload``: the gradient travels back through :func:`download`."""

    @staticmethod
    def forward(ctx, t, device):
        return upload(t, device)

    @staticmethod
    def backward(ctx, grad):
        return download(grad), None


def upload_with_grad(t, device=None):
    return _StagedUpload.apply(t, device)


__all__ = ["upload", "download", "upload_with_grad", "CHUNK_BYTES"]

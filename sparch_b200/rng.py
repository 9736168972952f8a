"""The reference's initial-state draws on the device.

``torch.rand(B, H)`` on the default CPU generator (sparch/models/snns.py:286-287, 423-425, 558-559, 700-702, 812) is a serial
MT19937 stream: ~1 ms of host time per (256, 1024) draw, seven draws per step of BASELINE cfg 4 -- the step of the
drop-in (``set_state_init("cpu")``, the default) was bound by it.  ``cpu_generator_rand`` returns the SAME numbers from a
device kernel (csrc/rng.cu) and leaves the CPU generator exactly where the host draws would have left it (its state
crosses PCIe: 2.5 KB each way), so everything else that uses the generator -- the DataLoader's sampler, parameter
initialisation -- sees the reference's stream.  Bit-exact by test against ``torch.rand`` (tests/test_gpu_parity.py).
"""
import struct

import numpy as np
import torch

from ._lib import call, ptr

_STATE_BYTES = 5056          # at::CPUGeneratorImpl's legacy state blob: seed u64, left i32, seeded i32, next u64,
_OFF_LEFT, _OFF_NEXT, _OFF_STATE, _MT_N = 8, 16, 24, 624   # state u64[624], normal-distribution cache
_streams = {}


def _rng_stream(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _streams:
        _streams[idx] = torch.cuda.Stream(device=idx, priority=-1)
    return _streams[idx]


_ahead = {}        # device index -> the draws prepared for the NEXT call (see cpu_generator_rand)


def _launch(words, pos, n, device, side, pinned):
    """n draws from (words, pos) on the stream ``side``; returns (out, host copy of [new state, new pos], event).  With
    ``pinned`` the copy back is asynchronous (the event tells when it is there), else the host waits for it."""
    st_in = torch.from_numpy(words.view(np.int32)).to(device)
    st_out = torch.empty(_MT_N + 1, dtype=torch.int32, device=device)
    out = torch.empty(n, dtype=torch.float32, device=device)
    call("sparch_mt19937_uniform", ptr(st_in), pos, n, ptr(out), ptr(st_out), side.cuda_stream)
    if pinned:
        host = torch.empty(_MT_N + 1, dtype=torch.int32, pin_memory=True)
        host.copy_(st_out, non_blocking=True)
    else:
        host = st_out.cpu()                                  # (waits for this stream only)
    done = torch.cuda.Event()
    done.record(side)
    return out, host, done


def cpu_generator_rand(n, device):
    """n float32 values on ``device``, identical to ``torch.rand(n)`` drawn NOW from the default CPU generator, which is
    advanced by n draws.  Returns a flat tensor (ready for use on the current stream), or None when the generator's state
    does not have the known layout (the caller then draws on the host as the reference does).

    A training loop asks for the same n every step, and between two steps nothing else may have touched the generator:
    after serving a call, the draws of the NEXT call are started at once from the state just set (one CTA on a
    high-priority stream, under the step in flight).  The next call uses them if the generator still holds exactly that
    state and n is the same -- no waiting for the kernel -- and computes on the spot otherwise; the numbers are the
    same either way."""
    blob = torch.get_rng_state()
    if blob.numel() != _STATE_BYTES or n <= 0:
        return None
    raw = bytearray(blob.numpy().tobytes())
    left, = struct.unpack_from("<i", raw, _OFF_LEFT)
    if not 1 <= left <= _MT_N + 1:
        return None
    device = torch.device(device)
    with torch.cuda.device(device):
        idx = torch.cuda.current_device()
        main, side = torch.cuda.current_stream(), _rng_stream(device)
        pre = _ahead.pop(idx, None)
        if pre is not None and pre["n"] == n and pre["before"] == bytes(raw):
            out, host, done = pre["out"], pre["host"], pre["done"]
            done.synchronize()                  # (long past: the kernel ran under the previous step)
        else:
            pos = _MT_N + 1 - left              # words of the current block already handed out (left = 1: all of them)
            words = np.frombuffer(raw, dtype=np.uint64, count=_MT_N, offset=_OFF_STATE).astype(np.uint32)
            with torch.cuda.stream(side):       # its own high-priority stream: not queued behind the step in flight
                out, host, done = _launch(words, pos, n, device, side, pinned=False)
        out.record_stream(main)
        main.wait_event(done)
        new = host.numpy().view(np.uint32)
        new_pos = int(new[_MT_N])
        struct.pack_into("<i", raw, _OFF_LEFT, _MT_N + 1 - new_pos)
        struct.pack_into("<Q", raw, _OFF_NEXT, new_pos)
        raw[_OFF_STATE:_OFF_STATE + 8 * _MT_N] = new[:_MT_N].astype(np.uint64).tobytes()
        torch.set_rng_state(torch.frombuffer(raw, dtype=torch.uint8).clone())
        with torch.cuda.stream(side):           # the next call's draws, from the state the generator holds now
            o2, h2, d2 = _launch(new[:_MT_N].copy(), new_pos, n, device, side, pinned=True)
        _ahead[idx] = {"n": n, "before": bytes(raw), "out": o2, "host": h2, "done": d2}
    return out

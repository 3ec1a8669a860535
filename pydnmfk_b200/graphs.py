"""CUDA-graph capture of one update step.

An MU / HALS iteration is a fixed sequence of ~15 kernel launches (plus 2-4 NCCL collectives on a grid); on small
shards (configs 1 and 5, or 65536^2 split over 8 GPUs where a pass takes < 0.5 ms) the Python + ctypes launch path
costs more than the kernels.  The step is therefore captured once into a CUDA graph and replayed: one graph for a plain
step and one for a step followed by the every-10th-iteration clamp (pyDNMF.py:155-157).

Requirements on the captured callable: no host synchronisation (`.item()`, `.cpu()`), no workspace growth (run it once
eagerly first) -- true for the FRO/KL multiplicative updates and the HALS sweeps; BCD (host-side accept/restore branch)
is never captured.
"""
import os

import torch
import torch.distributed as dist


def graphs_enabled(comm, method):
    """Capture when the step is graph-safe: not BCD, and collectives (if any) run over NCCL."""
    if os.environ.get('DNMF_NO_GRAPH'):
        return False
    if method.lower() == 'bcd':
        return False
    if comm is not None and comm.size > 1:
        return dist.is_initialized() and dist.get_backend() == 'nccl'
    return True


_side_streams = {}


def _capture_stream(device):
    """One capture stream per device (capture must not run on the legacy default stream)."""
    key = (device.type, device.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


class StepGraphs:
    """`plain()` / `clamped()` replay the captured step (+ clamp).  Capture happens lazily on first use; the callables
    must already have been run eagerly once (kernel attributes set, tf32 calibration done, workspace sized)."""

    def __init__(self, step_fn, clamp_fn):
        self._step, self._clamp = step_fn, clamp_fn
        self._g_plain = self._g_clamp = None
        self._pool = None

    def _capture(self, with_clamp):
        # Manual capture on a side stream.  `with torch.cuda.graph(g)` would also run gc.collect() and
        # torch.cuda.empty_cache() on entry: ~20 ms per capture, as much as a whole 1000-iteration fit of an NMFk
        # ensemble member costs on the device (two captures per fit; tools/prof_cfg5.py).
        g = torch.cuda.CUDAGraph()
        cur = torch.cuda.current_stream()
        side = _capture_stream(cur.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            if self._pool is not None:
                g.capture_begin(pool=self._pool)
            else:
                g.capture_begin()
            try:
                self._step()
                if with_clamp:
                    self._clamp()
            finally:
                g.capture_end()
        cur.wait_stream(side)
        if self._pool is None:
            self._pool = g.pool()
        return g

    def plain(self):
        if self._g_plain is None:
            self._g_plain = self._capture(False)
        self._g_plain.replay()

    def clamped(self):
        if self._g_clamp is None:
            self._g_clamp = self._capture(True)
        self._g_clamp.replay()

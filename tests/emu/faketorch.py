"""TEST INFRASTRUCTURE ONLY: the handful of torch names qpad_b200.pipeline.LocalPipeline touches (streams, events, device buffers),
for running it against the host emulation of the library: operations execute when they are enqueued, so streams and events are
bookkeeping only and a "device" tensor is a numpy array.  Installed as sys.modules["torch"] for the duration of a test."""
import contextlib
import os
import types

import numpy as np

float64 = np.float64


class _Tensor:
    def __init__(self, a): self.a = a
    def data_ptr(self): return self.a.ctypes.data
    def cpu(self): return self
    def numpy(self): return self.a.copy()
    def zero_(self): self.a[:] = 0; return self
    def item(self): return self.a.item()
    def clone(self): return _Tensor(self.a.copy())
    def __getitem__(self, k): return _Tensor(self.a[k])
    def __len__(self): return len(self.a)


def zeros(n, dtype=float64, device=None): return _Tensor(np.zeros(n, dtype=dtype))
def tensor(values, dtype=float64, device=None): return _Tensor(np.asarray(values, dtype=dtype))
def empty(n, dtype=float64, device=None): return zeros(n, dtype, device)
def device(kind, index=0): return (kind, index)


class _Stream:
    _next = 1

    def __init__(self, device=None):
        _Stream._next += 1
        self.cuda_stream = _Stream._next

    def wait_event(self, ev): assert ev.recorded, "wait on an event that was never recorded"
    def synchronize(self): pass


class _Event:
    def __init__(self, enable_timing=False): self.recorded = False
    def record(self, stream=None): self.recorded = True
    def synchronize(self): pass
    def elapsed_time(self, other): return 1.0          # not a clock: a non-zero constant so that rates can be formed


@contextlib.contextmanager
def _stream_ctx(s):
    yield


cuda = types.SimpleNamespace(Stream=_Stream, Event=_Event, stream=_stream_ctx, synchronize=lambda *a: None, is_available=lambda: True,
                             set_device=lambda d: None, current_stream=lambda *a: _Stream(),
                             get_device_properties=lambda d: types.SimpleNamespace(multi_processor_count=int(os.environ.get("QPAD_EMU_SMS", "12")), name="emulated"))

distributed = types.ModuleType("torch.distributed")          # world = 1 only: imported by bench.py, never called


def install(monkeypatch=None):
    """put this module in the place of torch (and torch.distributed) in sys.modules"""
    import sys
    me = sys.modules[__name__]
    for name, mod in (("torch", me), ("torch.distributed", distributed)):
        if monkeypatch is not None:
            monkeypatch.setitem(sys.modules, name, mod)
        else:
            sys.modules[name] = mod

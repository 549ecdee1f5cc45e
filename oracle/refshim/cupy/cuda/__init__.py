"""Stub of cupy.cuda for the NumPy-backed shim (test infrastructure only)."""
import sys
import types


class Device:
    def __init__(self, device=None):
        self.id = 0 if device is None else int(device)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def use(self):
        pass

    def synchronize(self):
        pass


class Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a, **k):
        pass

    def synchronize(self):
        pass


class Stream:
    null = None

    def __init__(self, *a, **k):
        self.ptr = 0

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def synchronize(self):
        pass

    def wait_event(self, e):
        pass

    def record(self, e=None):
        return e if e is not None else Event()

    def use(self):
        pass


Stream.null = Stream()


def get_current_stream():
    return Stream.null


runtime = types.ModuleType('cupy.cuda.runtime')
runtime.getDeviceCount = lambda: 1
runtime.getDevice = lambda: 0
runtime.setDevice = lambda i: None
runtime.deviceSynchronize = lambda: None

cufft = types.ModuleType('cupy.cuda.cufft')


class _Plan:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


cufft.Plan1d = _Plan
cufft.PlanNd = _Plan

memory = types.ModuleType('cupy.cuda.memory')


class OutOfMemoryError(MemoryError):
    pass


memory.OutOfMemoryError = OutOfMemoryError

profiler = types.ModuleType('cupy.cuda.profiler')
profiler.start = lambda: None
profiler.stop = lambda: None

for _m in (runtime, cufft, memory, profiler):
    sys.modules[_m.__name__] = _m

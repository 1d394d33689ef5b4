"""Device-resident buffers over the C ABI: keep data in HBM across calls so chains such as
rlft3 -> pointwise product -> rlft3^-1 never cross PCIe.  Thin RAII-style wrappers of nrb_device_alloc /
nrb_upload / nrb_download and the plan API; no torch needed."""
import numpy as np

from . import _lib


class DeviceArray:
    """`count` float64 values in device memory (zero-initialised)."""

    def __init__(self, lib, count):
        self.lib, self.count = lib, int(count)
        self.ptr = lib.device_alloc(8 * max(1, self.count))

    @classmethod
    def from_host(cls, lib, host, stream=0):
        host = np.ascontiguousarray(host, dtype=np.float64)
        d = cls(lib, host.size)
        lib.upload(d.ptr, host, stream)
        return d

    def to_host(self, out=None, stream=0):
        out = np.empty(self.count, dtype=np.float64) if out is None else out
        self.lib.download(out.reshape(-1), self.ptr, stream)
        self.lib.stream_synchronize(stream)
        return out

    def free(self):
        if self.ptr:
            self.lib.device_free(self.ptr)
            self.ptr = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()


class Rlft3Convolver:
    """3-D circular convolution with a fixed kernel, all on the device: NR's use of rlft3
    (forward both, multiply the spectra -- `data` and `speq` --, inverse, times 2/(nn1 nn2 nn3))."""

    def __init__(self, lib, kernel):
        self.lib = lib
        self.shape = tuple(kernel.shape)
        n1, n2, n3 = self.shape
        self.plan = lib.plan_create(_lib.KIND_RLFT3, [n1, n2, n3])
        self.kd = DeviceArray.from_host(lib, kernel)
        self.ks = DeviceArray(lib, 2 * n1 * n2)
        self.plan.exec(self.kd.ptr, self.ks.ptr, isign=1)
        self.xs = DeviceArray(lib, 2 * n1 * n2)

    def apply(self, x_dev):
        """x_dev: DeviceArray holding the volume; convolved in place, stays on the device."""
        n1, n2, n3 = self.shape
        self.plan.exec(x_dev.ptr, self.xs.ptr, isign=1)
        scale = 2.0 / (n1 * n2 * n3)
        self.lib.complex_multiply_device(x_dev.ptr, self.kd.ptr, n1 * n2 * n3 // 2, False, scale)
        self.lib.complex_multiply_device(self.xs.ptr, self.ks.ptr, n1 * n2, False, scale)
        self.plan.exec(x_dev.ptr, self.xs.ptr, isign=-1)
        return x_dev

    def close(self):
        self.plan.destroy()
        for d in (self.kd, self.ks, self.xs):
            d.free()

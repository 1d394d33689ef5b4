"""ctypes binding of the C ABI declared in include/numrs_b200.h.

`Library(path)` wraps one shared object exporting that ABI.  The product package binds
numrs_b200/libnumrs_b200.so (CUDA, sm_100a) and nothing else (see numrs_b200/__init__.py).
"""
import ctypes
import os

import numpy as np

_dp = ctypes.POINTER(ctypes.c_double)
_sz = ctypes.c_size_t
_vp = ctypes.c_void_p

# error codes of include/numrs_b200.h
NRB_OK = 0
NRB_ERR_EMPTY_INPUT = -1
NRB_ERR_RESPONSE_TOO_LONG = -2
NRB_ERR_INVALID_ISIGN = -3
NRB_ERR_LENGTH_MISMATCH = -4
NRB_ERR_INVALID_DIMS = -5
NRB_ERR_NOT_POW2 = -6
NRB_ERR_UNSUPPORTED = -7
NRB_ERR_ZERO_STDDEV = -8
NRB_ERR_CUDA = -10
NRB_ERR_NCCL = -11
NRB_ERR_OOM = -12

NRB_PAD_LITERAL = 0
NRB_PAD_NR = 1

KIND_FOUR1, KIND_FOURN, KIND_REALFT, KIND_RLFT3, KIND_CONVLV, KIND_CORREL = 1, 2, 3, 4, 5, 6
KIND_CORREL_NORM, KIND_CORREL_NORM_FAST, KIND_AUTOCORREL_FAST, KIND_TWOFFT, KIND_POWER = 7, 8, 9, 10, 11
KIND_COSFT1, KIND_COSFT2, KIND_SINFT = 12, 13, 14

# every symbol include/numrs_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "nrb_version", "nrb_last_error", "nrb_device_count", "nrb_set_device", "nrb_shutdown",
    "nrb_set_option", "nrb_host_alloc", "nrb_host_free", "nrb_num_devices_in_use", "nrb_multi_device_calls",
    "nrb_four1", "nrb_four1_batch", "nrb_fourn", "nrb_realft", "nrb_realft_batch", "nrb_rlft3",
    "nrb_convlv", "nrb_convlv_batch", "nrb_correl", "nrb_correl_batch",
    "nrb_correl_normalized", "nrb_autocorrel_fast", "nrb_twofft", "nrb_twofft_batch", "nrb_power_spectrum",
    "nrb_cosft1", "nrb_cosft2", "nrb_sinft",
    "nrb_plan_create", "nrb_plan_workspace_bytes", "nrb_plan_num_launches", "nrb_plan_exec", "nrb_plan_destroy",
    "nrb_plan_profile", "nrb_plan_describe_launch", "nrb_fill_uniform_device",
    "nrb_slab_create", "nrb_slab_create_fourn", "nrb_slab_exec", "nrb_slab_num_launches", "nrb_slab_local_doubles", "nrb_slab_speq_doubles", "nrb_slab_xchg_doubles",
    "nrb_slab_stage", "nrb_slab_destroy", "nrb_slab_set_peers", "nrb_slab_set_send_peers", "nrb_slab_recv_bytes", "nrb_slab_barrier",
    "nrb_slab_set_chunks", "nrb_slab_stage_part", "nrb_slab_barrier_chunk",
    "nrb_slab_set_dma", "nrb_slab_exec_dma", "nrb_slab_stage_part_xchg", "nrb_slab_dma_timeline",
    "nrb_upload", "nrb_download", "nrb_stream_synchronize", "nrb_complex_multiply_device",
    "nrb_device_alloc", "nrb_device_free", "nrb_ipc_export", "nrb_ipc_import", "nrb_ipc_release",
]


class NrbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"numrs_b200 error {code}: {message}")
        self.code = code
        self.message = message


def _f64(a):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
        raise TypeError("expected a C-contiguous float64 numpy array")
    return a.ctypes.data_as(_dp)


class Library:
    def __init__(self, path):
        if not os.path.exists(path):
            raise ImportError(
                f"{path} not found: the numrs_b200 CUDA library is not built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        self.path = path
        L = self.L = ctypes.CDLL(path)
        L.nrb_version.restype = ctypes.c_char_p
        L.nrb_last_error.restype = ctypes.c_char_p
        L.nrb_device_count.restype = ctypes.c_int
        L.nrb_set_device.argtypes = [ctypes.c_int]
        L.nrb_set_option.argtypes = [ctypes.c_char_p, ctypes.c_long]
        L.nrb_host_alloc.argtypes = [_sz]
        L.nrb_host_alloc.restype = _vp
        L.nrb_host_free.argtypes = [_vp]
        L.nrb_host_free.restype = None
        L.nrb_four1.argtypes = [_dp, _sz, ctypes.c_int]
        L.nrb_four1_batch.argtypes = [ctypes.POINTER(_dp), ctypes.POINTER(_sz), _sz, ctypes.c_int]
        L.nrb_fourn.argtypes = [_dp, ctypes.POINTER(_sz), _sz, ctypes.c_int]
        L.nrb_realft.argtypes = [_dp, _sz, ctypes.c_int]
        L.nrb_realft_batch.argtypes = [ctypes.POINTER(_dp), _sz, _sz, ctypes.c_int]
        L.nrb_rlft3.argtypes = [_dp, _dp, _sz, _sz, _sz, ctypes.c_int]
        L.nrb_convlv.argtypes = [_dp, _sz, _dp, _sz, ctypes.c_int, ctypes.c_int, _dp]
        L.nrb_convlv_batch.argtypes = [ctypes.POINTER(_dp), _sz, _sz, _dp, _sz, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(_dp)]
        L.nrb_correl.argtypes = [_dp, _sz, _dp, _sz, _dp]
        L.nrb_correl_batch.argtypes = [ctypes.POINTER(_dp), ctypes.POINTER(_dp), _sz, _sz, ctypes.POINTER(_dp)]
        L.nrb_correl_normalized.argtypes = [_dp, _sz, _dp, _sz, ctypes.c_int, _dp]
        L.nrb_autocorrel_fast.argtypes = [_dp, _sz, _dp]
        L.nrb_twofft.argtypes = [_dp, _dp, _sz, _dp, _dp]
        L.nrb_twofft_batch.argtypes = [ctypes.POINTER(_dp), ctypes.POINTER(_dp), _sz, _sz, ctypes.POINTER(_dp), ctypes.POINTER(_dp)]
        L.nrb_power_spectrum.argtypes = [_dp, _sz, ctypes.c_int, _dp]
        L.nrb_cosft1.argtypes = [_dp, _sz]
        L.nrb_cosft2.argtypes = [_dp, _sz, ctypes.c_int]
        L.nrb_sinft.argtypes = [_dp, _sz]
        L.nrb_plan_create.argtypes = [ctypes.c_int, ctypes.POINTER(_sz), _sz, _sz, ctypes.POINTER(_vp)]
        L.nrb_plan_workspace_bytes.argtypes = [_vp]
        L.nrb_plan_workspace_bytes.restype = _sz
        L.nrb_plan_num_launches.argtypes = [_vp, ctypes.c_int]
        L.nrb_plan_exec.argtypes = [_vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp]
        L.nrb_plan_destroy.argtypes = [_vp]
        L.nrb_plan_profile.argtypes = [_vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp,
                                       ctypes.POINTER(ctypes.c_float), ctypes.c_int]
        L.nrb_plan_describe_launch.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, _sz, _dp]
        L.nrb_fill_uniform_device.argtypes = [_vp, ctypes.c_ulonglong, ctypes.c_ulonglong, _sz, _vp]
        L.nrb_slab_create.argtypes = [_sz, _sz, _sz, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)]
        L.nrb_multi_device_calls.argtypes = [ctypes.c_int]
        L.nrb_multi_device_calls.restype = ctypes.c_long
        L.nrb_slab_create_fourn.argtypes = [_sz, _sz, _sz, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)]
        L.nrb_slab_num_launches.argtypes = [_vp, ctypes.c_int]
        L.nrb_slab_exec.argtypes = [_vp, ctypes.c_int, _vp, _vp, ctypes.c_ulonglong, _vp]
        for n in ("nrb_slab_local_doubles", "nrb_slab_speq_doubles", "nrb_slab_xchg_doubles"):
            getattr(L, n).argtypes = [_vp]
            getattr(L, n).restype = _sz
        L.nrb_slab_stage.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]
        L.nrb_slab_destroy.argtypes = [_vp]
        L.nrb_slab_set_peers.argtypes = [_vp, ctypes.POINTER(_vp), ctypes.c_int]
        L.nrb_slab_set_send_peers.argtypes = [_vp, ctypes.POINTER(_vp), ctypes.c_int]
        L.nrb_slab_recv_bytes.argtypes = [_vp]
        L.nrb_slab_recv_bytes.restype = _sz
        L.nrb_slab_barrier.argtypes = [_vp, ctypes.c_int, ctypes.c_ulonglong, _vp]
        L.nrb_slab_set_chunks.argtypes = [_vp, ctypes.c_int]
        L.nrb_slab_stage_part.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]
        L.nrb_slab_barrier_chunk.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong, _vp]
        L.nrb_upload.argtypes = [_vp, _vp, _sz, _vp]
        L.nrb_download.argtypes = [_vp, _vp, _sz, _vp]
        L.nrb_stream_synchronize.argtypes = [_vp]
        L.nrb_complex_multiply_device.argtypes = [_vp, _vp, _sz, ctypes.c_int, ctypes.c_double, _vp]
        L.nrb_slab_set_dma.argtypes = [_vp, ctypes.c_int]
        L.nrb_slab_exec_dma.argtypes = [_vp, ctypes.c_int, _vp, _vp, ctypes.c_ulonglong, _vp]
        L.nrb_slab_stage_part_xchg.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp]
        L.nrb_slab_dma_timeline.argtypes = [_vp, ctypes.c_int, ctypes.c_char_p, _sz]
        L.nrb_device_alloc.argtypes = [_sz, ctypes.POINTER(_vp)]
        L.nrb_device_free.argtypes = [_vp]
        L.nrb_ipc_export.argtypes = [_vp, ctypes.c_char_p]
        L.nrb_ipc_import.argtypes = [ctypes.c_char_p, ctypes.POINTER(_vp)]
        L.nrb_ipc_release.argtypes = [_vp]

    # ---- helpers
    def last_error(self):
        return self.L.nrb_last_error().decode()

    def check(self, rc):
        if rc != 0:
            raise NrbError(rc, self.last_error())

    def version(self):
        return self.L.nrb_version().decode()

    def device_count(self):
        return self.L.nrb_device_count()

    def set_device(self, dev):
        self.check(self.L.nrb_set_device(dev))

    def set_option(self, name, value):
        self.check(self.L.nrb_set_option(name.encode(), int(value)))

    def shutdown(self):
        self.L.nrb_shutdown()

    def pinned_empty(self, count):
        """float64 numpy array of `count` elements in pinned host memory (nrb_host_alloc).
        The memory stays allocated until pinned_free(arr) or process exit."""
        p = self.L.nrb_host_alloc(max(8, count * 8))
        if not p:
            raise MemoryError("nrb_host_alloc failed")
        buf = (ctypes.c_double * count).from_address(p)
        arr = np.frombuffer(buf, dtype=np.float64, count=count)
        _PINNED[arr.ctypes.data] = p
        return arr

    def pinned_free(self, arr):
        p = _PINNED.pop(arr.ctypes.data, None)
        if p:
            self.L.nrb_host_free(p)

    # ---- raw host-slice calls (return the C return code)
    def four1(self, data, nn, isign):
        return self.L.nrb_four1(_f64(data), nn, isign)

    def four1_batch(self, arrays, isign, nn=None):
        cnt = len(arrays)
        ptrs = (_dp * max(cnt, 1))(*[_f64(a) for a in arrays])
        sizes = (_sz * max(cnt, 1))(*([a.size // 2 for a in arrays] if nn is None else nn))
        return self.L.nrb_four1_batch(ptrs, sizes, cnt, isign)

    def batch_table(self, arrays):
        """Pointer / length tables of a four1 batch, built once and reusable across calls (a Rust caller's Vec of
        pointers costs nothing; 4096 ctypes conversions per call cost more than the PCIe copies)."""
        cnt = len(arrays)
        ptrs = (_dp * max(cnt, 1))(*[_f64(a) for a in arrays])
        sizes = (_sz * max(cnt, 1))(*[a.size // 2 for a in arrays])
        return (ptrs, sizes, cnt, list(arrays))      # the arrays stay referenced

    def four1_batch_table(self, table, isign):
        ptrs, sizes, cnt, _ = table
        return self.L.nrb_four1_batch(ptrs, sizes, cnt, isign)

    def fourn(self, data, nn, ndim, isign):
        nn_c = (_sz * max(len(nn), 1))(*nn)
        return self.L.nrb_fourn(_f64(data), nn_c, ndim, isign)

    def realft(self, data, n, isign):
        return self.L.nrb_realft(_f64(data), n, isign)

    def realft_batch(self, arrays, n, isign):
        cnt = len(arrays)
        ptrs = (_dp * max(cnt, 1))(*[_f64(a) for a in arrays])
        return self.L.nrb_realft_batch(ptrs, n, cnt, isign)

    def rlft3(self, data, speq, nn1, nn2, nn3, isign):
        return self.L.nrb_rlft3(_f64(data), _f64(speq), nn1, nn2, nn3, isign)

    def convlv(self, data, respns, isign, pad_mode=NRB_PAD_LITERAL):
        data = np.ascontiguousarray(data, dtype=np.float64)
        respns = np.ascontiguousarray(respns, dtype=np.float64)
        ans = np.zeros(max(1, data.size), dtype=np.float64)
        d = data if data.size else np.zeros(1)
        r = respns if respns.size else np.zeros(1)
        rc = self.L.nrb_convlv(_f64(d), data.size, _f64(r), respns.size, isign, pad_mode, _f64(ans))
        return rc, ans[:data.size]

    def convlv_batch(self, signals, respns, isign, pad_mode=NRB_PAD_LITERAL, outs=None):
        cnt = len(signals)
        n = signals[0].size if cnt else 0
        respns = np.ascontiguousarray(respns, dtype=np.float64)
        if outs is None:
            outs = [np.zeros(n, dtype=np.float64) for _ in range(cnt)]
        ip = (_dp * max(cnt, 1))(*[_f64(s) for s in signals])
        op = (_dp * max(cnt, 1))(*[_f64(o) for o in outs])
        rc = self.L.nrb_convlv_batch(ip, cnt, n, _f64(respns), respns.size, isign, pad_mode, op)
        return rc, outs

    def correl(self, d1, d2):
        d1 = np.ascontiguousarray(d1, dtype=np.float64)
        d2 = np.ascontiguousarray(d2, dtype=np.float64)
        ans = np.zeros(max(1, d1.size), dtype=np.float64)
        a = d1 if d1.size else np.zeros(1)
        b = d2 if d2.size else np.zeros(1)
        rc = self.L.nrb_correl(_f64(a), d1.size, _f64(b), d2.size, _f64(ans))
        return rc, ans[:d1.size]

    def correl_batch(self, a_list, b_list, outs=None):
        cnt = len(a_list)
        n = a_list[0].size if cnt else 0
        if outs is None:
            outs = [np.zeros(n, dtype=np.float64) for _ in range(cnt)]
        ap = (_dp * max(cnt, 1))(*[_f64(a) for a in a_list])
        bp = (_dp * max(cnt, 1))(*[_f64(b) for b in b_list])
        op = (_dp * max(cnt, 1))(*[_f64(o) for o in outs])
        rc = self.L.nrb_correl_batch(ap, bp, cnt, n, op)
        return rc, outs

    def correl_normalized(self, d1, d2, fast=False):
        d1 = np.ascontiguousarray(d1, dtype=np.float64)
        d2 = np.ascontiguousarray(d2, dtype=np.float64)
        ans = np.zeros(max(1, d1.size), dtype=np.float64)
        a = d1 if d1.size else np.zeros(1)
        b = d2 if d2.size else np.zeros(1)
        rc = self.L.nrb_correl_normalized(_f64(a), d1.size, _f64(b), d2.size, int(fast), _f64(ans))
        return rc, ans[:d1.size]

    def autocorrel_fast(self, d):
        d = np.ascontiguousarray(d, dtype=np.float64)
        ans = np.zeros(max(1, d.size), dtype=np.float64)
        rc = self.L.nrb_autocorrel_fast(_f64(d if d.size else np.zeros(1)), d.size, _f64(ans))
        return rc, ans[:d.size]

    def twofft(self, d1, d2, fft1, fft2):
        return self.L.nrb_twofft(_f64(d1), _f64(d2), d1.size, _f64(fft1), _f64(fft2))

    def twofft_batch(self, d1_list, d2_list, fft1_list, fft2_list):
        cnt = len(d1_list)
        n = d1_list[0].size if cnt else 1
        arr = lambda xs: (_dp * max(cnt, 1))(*[_f64(x) for x in xs])   # noqa: E731
        return self.L.nrb_twofft_batch(arr(d1_list), arr(d2_list), cnt, n, arr(fft1_list), arr(fft2_list))

    def power_spectrum(self, c, take_sqrt=False):
        c = np.ascontiguousarray(c, dtype=np.float64)
        out = np.zeros(c.size // 2, dtype=np.float64)
        rc = self.L.nrb_power_spectrum(_f64(c if c.size else np.zeros(2)), c.size // 2, int(take_sqrt),
                                       _f64(out if out.size else np.zeros(1)))
        return rc, out

    def cosft1(self, y, n):
        return self.L.nrb_cosft1(_f64(y), n)

    def cosft2(self, y, n, isign):
        return self.L.nrb_cosft2(_f64(y), n, isign)

    def sinft(self, y, n):
        return self.L.nrb_sinft(_f64(y), n)

    def fill_uniform_device(self, d_ptr, seed, offset, count, stream=0):
        self.check(self.L.nrb_fill_uniform_device(d_ptr, seed, offset, count, stream or None))

    def device_alloc(self, nbytes):
        p = _vp()
        self.check(self.L.nrb_device_alloc(nbytes, ctypes.byref(p)))
        return p.value

    def device_free(self, ptr):
        self.L.nrb_device_free(ptr)

    def upload(self, d_ptr, host, stream=0):
        host = np.ascontiguousarray(host, dtype=np.float64)
        self.check(self.L.nrb_upload(d_ptr, host.ctypes.data, host.nbytes, stream or None))

    def download(self, host, d_ptr, stream=0):
        assert host.dtype == np.float64 and host.flags["C_CONTIGUOUS"]
        self.check(self.L.nrb_download(host.ctypes.data, d_ptr, host.nbytes, stream or None))

    def stream_synchronize(self, stream=0):
        self.check(self.L.nrb_stream_synchronize(stream or None))

    def complex_multiply_device(self, d_a, d_b, ncomplex, conj_b=False, scale=1.0, stream=0):
        self.check(self.L.nrb_complex_multiply_device(d_a, d_b, ncomplex, int(conj_b), float(scale), stream or None))

    def ipc_export(self, ptr):
        buf = ctypes.create_string_buffer(64)
        self.check(self.L.nrb_ipc_export(ptr, buf))
        return buf.raw

    def ipc_import(self, handle):
        p = _vp()
        self.check(self.L.nrb_ipc_import(handle, ctypes.byref(p)))
        return p.value

    def ipc_release(self, ptr):
        self.check(self.L.nrb_ipc_release(ptr))

    # ---- device-resident plan API (pointers are integers: device addresses)
    def plan_create(self, kind, dims, batch=1):
        h = _vp()
        dims_c = (_sz * max(len(dims), 1))(*dims)
        self.check(self.L.nrb_plan_create(kind, dims_c, len(dims), batch, ctypes.byref(h)))
        return Plan(self, h)

    def num_devices_in_use(self):
        return self.L.nrb_num_devices_in_use()

    def multi_device_calls(self, which=0):
        return self.L.nrb_multi_device_calls(which)

    def slab_create(self, nn1, nn2, nn3, nranks, rank, kind="rlft3"):
        """kind "rlft3" (real volume + speq plane) or "fourn" (3-D complex volume of nn3 complex points per line)."""
        h = _vp()
        create = self.L.nrb_slab_create if kind == "rlft3" else self.L.nrb_slab_create_fourn
        self.check(create(nn1, nn2, nn3, nranks, rank, ctypes.byref(h)))
        return SlabPlan(self, h)


_PINNED = {}


class Plan:
    def __init__(self, lib, handle):
        self.lib, self.h = lib, handle

    def workspace_bytes(self):
        return self.lib.L.nrb_plan_workspace_bytes(self.h)

    def num_launches(self, isign):
        return self.lib.L.nrb_plan_num_launches(self.h, isign)

    def exec(self, d_io, d_aux=0, d_out=0, isign=1, arg=0, stream=0):
        self.lib.check(self.lib.L.nrb_plan_exec(self.h, d_io, d_aux or None, d_out or None, isign, arg,
                                                stream or None))

    def profile(self, d_io, d_aux=0, d_out=0, isign=1, arg=0, stream=0):
        """Run once with events around every launch; returns [(name, algorithmic_bytes, ms)]."""
        n = self.num_launches(isign)
        ms = (ctypes.c_float * max(n, 1))()
        self.lib.check(self.lib.L.nrb_plan_profile(self.h, d_io, d_aux or None, d_out or None, isign, arg,
                                                   stream or None, ms, n))
        out = []
        for i in range(n):
            name = ctypes.create_string_buffer(128)
            b = ctypes.c_double()
            self.lib.L.nrb_plan_describe_launch(self.h, isign, i, name, 128, ctypes.byref(b))
            out.append((name.value.decode(), b.value, float(ms[i])))
        return out

    def destroy(self):
        if self.h:
            self.lib.L.nrb_plan_destroy(self.h)
            self.h = None


class SlabPlan:
    def __init__(self, lib, handle):
        self.lib, self.h = lib, handle

    def local_doubles(self):
        return self.lib.L.nrb_slab_local_doubles(self.h)

    def speq_doubles(self):
        return self.lib.L.nrb_slab_speq_doubles(self.h)

    def xchg_doubles(self):
        return self.lib.L.nrb_slab_xchg_doubles(self.h)

    def set_peers(self, peer_ptrs):
        """Fused exchange: peer_ptrs[i] = rank i's receive buffer mapped into this process (None to disable)."""
        if peer_ptrs is None:
            self.lib.check(self.lib.L.nrb_slab_set_peers(self.h, None, 0))
            return
        arr = (_vp * len(peer_ptrs))(*peer_ptrs)
        self.lib.check(self.lib.L.nrb_slab_set_peers(self.h, arr, len(peer_ptrs)))

    def set_send_peers(self, peer_ptrs):
        """Push + pull exchange: peer_ptrs[i] = rank i's send buffer mapped into this process (None to disable)."""
        if peer_ptrs is None:
            self.lib.check(self.lib.L.nrb_slab_set_send_peers(self.h, None, 0))
            return
        arr = (_vp * len(peer_ptrs))(*peer_ptrs)
        self.lib.check(self.lib.L.nrb_slab_set_send_peers(self.h, arr, len(peer_ptrs)))

    def recv_bytes(self):
        return self.lib.L.nrb_slab_recv_bytes(self.h)

    def barrier(self, phase, epoch, stream=0):
        self.lib.check(self.lib.L.nrb_slab_barrier(self.h, phase, epoch, stream or None))

    def exec(self, isign, d_slab, d_speq, epoch, stream=0):
        """One whole direction of the fused exchange (stage 0, flag barrier, stage 1) in one C call."""
        self.lib.check(self.lib.L.nrb_slab_exec(self.h, isign, d_slab, d_speq or None, epoch, stream or None))

    def num_launches(self, isign):
        return self.lib.L.nrb_slab_num_launches(self.h, isign)

    def set_chunks(self, chunks):
        self.lib.check(self.lib.L.nrb_slab_set_chunks(self.h, chunks))

    def stage_part(self, stage, part, isign, d_slab, d_speq, stream=0):
        self.lib.check(self.lib.L.nrb_slab_stage_part(self.h, stage, part, isign, d_slab, d_speq, stream or None))

    def set_dma(self, chunks):
        self.lib.check(self.lib.L.nrb_slab_set_dma(self.h, chunks))

    def exec_dma(self, isign, d_slab, d_speq, epoch, stream=0):
        self.lib.check(self.lib.L.nrb_slab_exec_dma(self.h, isign, d_slab, d_speq, epoch, stream or None))

    def stage_part_xchg(self, stage, part, isign, d_slab, d_speq, d_xchg, stream=0):
        self.lib.check(self.lib.L.nrb_slab_stage_part_xchg(self.h, stage, part, isign, d_slab, d_speq, d_xchg or None,
                                                           stream or None))

    def dma_timeline(self, enable):
        buf = ctypes.create_string_buffer(4096)
        self.lib.check(self.lib.L.nrb_slab_dma_timeline(self.h, int(enable), buf, 4096))
        return buf.value.decode()

    def barrier_chunk(self, phase, chunk, epoch, stream=0):
        self.lib.check(self.lib.L.nrb_slab_barrier_chunk(self.h, phase, chunk, epoch, stream or None))

    def stage(self, stage, isign, d_slab, d_speq, d_send, d_recv, stream=0):
        self.lib.check(self.lib.L.nrb_slab_stage(self.h, stage, isign, d_slab, d_speq, d_send or None,
                                                 d_recv or None, stream or None))

    def destroy(self):
        if self.h:
            self.lib.L.nrb_slab_destroy(self.h)
            self.h = None

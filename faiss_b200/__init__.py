"""faiss_b200 -- B200-native (sm_100a) similarity search behind the Faiss GPU plugin surface.

Python mirror of the reference's index classes over the C ABI of ``libfaiss_b200.so``
(``include/faiss_b200_c.h``).  Same class / method names and argument meaning as
``faiss.GpuIndexFlat{,L2,IP}``, ``faiss.GpuIndexIVFFlat``, ``faiss.GpuIndexIVFPQ``,
``faiss.StandardGpuResources``, ``faiss.IndexShards`` (faiss/python/gpu_wrappers.py:21-60,
faiss/gpu/GpuIndex*.h).  Inputs may be numpy arrays (host) or torch CUDA tensors (device);
outputs follow the input's residency.

There is no CPU fallback: importing this package without the compiled CUDA library fails.
"""
import ctypes
import json
import os

import numpy as np

from ._lib import lib, check, FaissError  # noqa: F401  (loading fails loudly)

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1

_c_f = ctypes.POINTER(ctypes.c_float)
_c_i64 = ctypes.POINTER(ctypes.c_int64)
_c_u8 = ctypes.POINTER(ctypes.c_uint8)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x, ctype):
    """Raw pointer of a numpy array or torch tensor (host or device)."""
    if x is None:
        return ctypes.cast(None, ctype)
    if _is_torch(x):
        assert x.is_contiguous(), "tensor must be contiguous"
        return ctypes.cast(x.data_ptr(), ctype)
    assert x.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return x.ctypes.data_as(ctype)


def _as_f32(x):
    if _is_torch(x):
        import torch

        assert x.dtype == torch.float32
        return x.contiguous()
    return np.ascontiguousarray(x, dtype=np.float32)


def _as_i64(x):
    if x is None:
        return None
    if _is_torch(x):
        import torch

        assert x.dtype == torch.int64
        return x.contiguous()
    return np.ascontiguousarray(x, dtype=np.int64)


def _empty_like_residency(x, shape, dtype):
    if _is_torch(x) and x.is_cuda:
        import torch

        tdt = {np.float32: torch.float32, np.int64: torch.int64}[dtype]
        return torch.empty(shape, dtype=tdt, device=x.device)
    return np.empty(shape, dtype=dtype)


class StandardGpuResources:
    """faiss::gpu::StandardGpuResources (faiss/gpu/StandardGpuResources.h:199-266)."""

    def __init__(self):
        self._h = ctypes.c_void_p()
        check(lib.faiss_StandardGpuResources_new(ctypes.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.faiss_StandardGpuResources_free(self._h)
            self._h = None

    def noTempMemory(self):
        check(lib.faiss_StandardGpuResources_noTempMemory(self._h))

    def setTempMemory(self, size):
        check(lib.faiss_StandardGpuResources_setTempMemory(self._h, ctypes.c_size_t(size)))

    def setPinnedMemory(self, size):
        check(lib.faiss_StandardGpuResources_setPinnedMemory(self._h, ctypes.c_size_t(size)))

    def setDefaultStream(self, device, stream):
        check(lib.faiss_StandardGpuResources_setDefaultStream(self._h, int(device), ctypes.c_void_p(int(stream))))

    def setDefaultNullStreamAllDevices(self):
        check(lib.faiss_StandardGpuResources_setDefaultNullStreamAllDevices(self._h))

    def getDefaultStream(self, device):
        out = ctypes.c_void_p()
        check(lib.faiss_StandardGpuResources_getDefaultStream(self._h, int(device), ctypes.byref(out)))
        return out.value or 0

    def syncDefaultStream(self, device):
        check(lib.faiss_StandardGpuResources_syncDefaultStream(self._h, int(device)))

    def getMemoryInfo(self):
        buf = ctypes.create_string_buffer(1 << 16)
        check(lib.faiss_StandardGpuResources_getMemoryInfo(self._h, buf, ctypes.c_size_t(len(buf))))
        raw = json.loads(buf.value.decode())
        return {int(d): {k: tuple(v) for k, v in m.items()} for d, m in raw.items()}

    # -- NCCL communicator ownership
    def ncclInitAll(self, devices):
        """all listed devices of this process in one clique (rank i = devices[i])"""
        arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        check(lib.faiss_StandardGpuResources_ncclInitAll(self._h, len(devices), arr))

    def ncclInitRank(self, device, nranks, rank, unique_id):
        """this process = rank `rank` of `nranks`; unique_id: the 128 bytes of nccl_unique_id() from one rank"""
        unique_id = bytes(unique_id)
        assert len(unique_id) == 128
        check(lib.faiss_StandardGpuResources_ncclInitRank(self._h, int(device), int(nranks), int(rank), unique_id))

    def ncclRank(self, device):
        r, n = ctypes.c_int(), ctypes.c_int()
        check(lib.faiss_StandardGpuResources_ncclRank(self._h, int(device), ctypes.byref(r), ctypes.byref(n)))
        return r.value, n.value

    def getTempMemoryAvailable(self, device):
        out = ctypes.c_size_t()
        check(lib.faiss_StandardGpuResources_getTempMemoryAvailable(self._h, int(device), ctypes.byref(out)))
        return out.value


class Index:
    """faiss::Index surface (faiss/Index.h:101-435) over an opaque C handle."""

    def __init__(self):
        self._h = ctypes.c_void_p()
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.faiss_Index_free(self._h)
            self._h = None

    def _use_torch_stream(self, *tensors):
        """PyTorch interop as in faiss.contrib.torch_utils: when an argument is a CUDA tensor, order
        the library's work on torch's current stream (StandardGpuResources::setDefaultStream,
        faiss/gpu/StandardGpuResources.h:75,232) so that torch ops before/after the call are
        correctly ordered with it."""
        for t in tensors:
            if t is not None and _is_torch(t) and t.is_cuda:
                import torch

                dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
                for r in self._resources():
                    r.setDefaultStream(dev, torch.cuda.current_stream(dev).cuda_stream)
                return

    def _resources(self):
        return [r for r in self._keep if isinstance(r, StandardGpuResources)]

    # -- fields
    @property
    def d(self):
        return lib.faiss_Index_d(self._h)

    @property
    def ntotal(self):
        return lib.faiss_Index_ntotal(self._h)

    @property
    def is_trained(self):
        return bool(lib.faiss_Index_is_trained(self._h))

    @property
    def metric_type(self):
        return lib.faiss_Index_metric_type(self._h)

    @property
    def verbose(self):
        return bool(lib.faiss_Index_verbose(self._h))

    @verbose.setter
    def verbose(self, v):
        lib.faiss_Index_set_verbose(self._h, int(bool(v)))

    # -- methods
    def _check_x(self, x):
        x = _as_f32(x)
        assert x.ndim == 2 and x.shape[1] == self.d, "x must be [n, d=%d], got %s" % (self.d, tuple(x.shape))
        return x

    def train(self, x):
        x = self._check_x(x)
        self._use_torch_stream(x)
        check(lib.faiss_Index_train(self._h, ctypes.c_int64(x.shape[0]), _ptr(x, _c_f)))

    def add(self, x):
        x = self._check_x(x)
        self._use_torch_stream(x)
        check(lib.faiss_Index_add(self._h, ctypes.c_int64(x.shape[0]), _ptr(x, _c_f)))

    def add_with_ids(self, x, ids):
        x = self._check_x(x)
        ids = _as_i64(ids)
        assert ids.shape == (x.shape[0],)
        self._use_torch_stream(x, ids)
        check(lib.faiss_Index_add_with_ids(self._h, ctypes.c_int64(x.shape[0]), _ptr(x, _c_f), _ptr(ids, _c_i64)))

    def setMinPagingSize(self, size):
        """GpuIndex::setMinPagingSize (faiss/gpu/GpuIndex.h:66-69)"""
        check(lib.faiss_GpuIndex_setMinPagingSize(self._h, ctypes.c_size_t(size)))

    def getMinPagingSize(self):
        out = ctypes.c_size_t()
        check(lib.faiss_GpuIndex_getMinPagingSize(self._h, ctypes.byref(out)))
        return int(out.value)

    def search(self, x, k, D=None, I=None, params=None):
        """params: a SearchParametersIVF (per-call nprobe) or None -- faiss::Index::search(..., params)"""
        x = self._check_x(x)
        n = x.shape[0]
        if D is None:
            D = _empty_like_residency(x, (n, k), np.float32)
        if I is None:
            I = _empty_like_residency(x, (n, k), np.int64)
        self._use_torch_stream(x, D, I)
        if params is not None:
            check(
                lib.faiss_Index_search_with_params(
                    self._h, ctypes.c_int64(n), _ptr(x, _c_f), ctypes.c_int64(k), params._h, _ptr(D, _c_f), _ptr(I, _c_i64)
                )
            )
            return D, I
        check(
            lib.faiss_Index_search(
                self._h, ctypes.c_int64(n), _ptr(x, _c_f), ctypes.c_int64(k), _ptr(D, _c_f), _ptr(I, _c_i64)
            )
        )
        return D, I

    def assign(self, x, k=1):
        x = self._check_x(x)
        n = x.shape[0]
        I = _empty_like_residency(x, (n, k), np.int64)
        self._use_torch_stream(x)
        check(lib.faiss_Index_assign(self._h, ctypes.c_int64(n), _ptr(x, _c_f), _ptr(I, _c_i64), ctypes.c_int64(k)))
        return I

    def reset(self):
        check(lib.faiss_Index_reset(self._h))

    def reconstruct(self, key):
        out = np.empty(self.d, dtype=np.float32)
        check(lib.faiss_Index_reconstruct(self._h, ctypes.c_int64(key), _ptr(out, _c_f)))
        return out

    def reconstruct_n(self, i0=0, ni=-1):
        if ni < 0:
            ni = self.ntotal - i0
        out = np.empty((ni, self.d), dtype=np.float32)
        check(lib.faiss_Index_reconstruct_n(self._h, ctypes.c_int64(i0), ctypes.c_int64(ni), _ptr(out, _c_f)))
        return out

    def reconstruct_batch(self, keys):
        keys = _as_i64(np.asarray(keys))
        out = np.empty((keys.shape[0], self.d), dtype=np.float32)
        check(lib.faiss_Index_reconstruct_batch(self._h, ctypes.c_int64(keys.shape[0]), _ptr(keys, _c_i64), _ptr(out, _c_f)))
        return out

    def compute_residual(self, x, key):
        x = _as_f32(x).reshape(1, -1)
        return self.compute_residual_n(x, np.array([key], dtype=np.int64))[0]

    def compute_residual_n(self, x, keys):
        x = self._check_x(x)
        keys = _as_i64(keys)
        out = _empty_like_residency(x, tuple(x.shape), np.float32)
        self._use_torch_stream(x, keys)
        check(
            lib.faiss_Index_compute_residual_n(
                self._h, ctypes.c_int64(x.shape[0]), _ptr(x, _c_f), _ptr(out, _c_f), _ptr(keys, _c_i64)
            )
        )
        return out


class SearchParametersIVF:
    """faiss::SearchParametersIVF (faiss/IndexIVF.h:68-90): per-call nprobe; max_codes must stay 0 on the GPU."""

    def __init__(self, nprobe=1, max_codes=0):
        self._h = ctypes.c_void_p()
        check(lib.faiss_SearchParametersIVF_new_with(ctypes.byref(self._h), ctypes.c_size_t(int(nprobe)), ctypes.c_size_t(int(max_codes))))

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.faiss_SearchParameters_free(self._h)
            self._h = None


_INTERRUPT_CB = None


def set_interrupt_callback(fn):
    """faiss::InterruptCallback: fn() -> truthy to make the running call fail with 'computation interrupted'; None clears."""
    global _INTERRUPT_CB
    proto = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p)
    if fn is None:
        lib.faiss_b200_set_interrupt_callback(ctypes.cast(None, proto), None)
        _INTERRUPT_CB = None
        return
    _INTERRUPT_CB = proto(lambda _ctx: 1 if fn() else 0)  # keep the thunk alive
    lib.faiss_b200_set_interrupt_callback(_INTERRUPT_CB, None)


class GpuIndexFlat(Index):
    """faiss::gpu::GpuIndexFlat (faiss/gpu/GpuIndexFlat.h:43-141)."""

    def __init__(self, res, d, metric=METRIC_L2, device=0, use_tensor_cores=True, use_float16=False):
        super().__init__()
        self._keep.append(res)
        check(
            lib.faiss_GpuIndexFlat_new_with_config(
                ctypes.byref(self._h), res._h, int(d), int(metric), int(device), int(bool(use_tensor_cores)),
                int(bool(use_float16)),
            )
        )

    def copyFrom(self, xb):
        xb = _as_f32(xb)
        check(lib.faiss_GpuIndexFlat_copyFrom(self._h, ctypes.c_int64(xb.shape[0]), _ptr(xb, _c_f)))

    def copyTo(self):
        out = np.empty((self.ntotal, self.d), dtype=np.float32)
        check(lib.faiss_GpuIndexFlat_copyTo(self._h, _ptr(out, _c_f)))
        return out

    def setUseTensorCores(self, enable):
        check(lib.faiss_GpuIndexFlat_setUseTensorCores(self._h, int(bool(enable))))

    def lastSearchInfo(self):
        out = (ctypes.c_int * 2)()
        check(lib.faiss_GpuIndexFlat_lastSearchInfo(self._h, out))
        return {"tensor_cores": int(out[0]), "fallback_queries": int(out[1])}


class GpuIndexFlatL2(GpuIndexFlat):
    def __init__(self, res, d, device=0, use_tensor_cores=True, use_float16=False):
        super().__init__(res, d, METRIC_L2, device, use_tensor_cores, use_float16)


class GpuIndexFlatIP(GpuIndexFlat):
    def __init__(self, res, d, device=0, use_tensor_cores=True, use_float16=False):
        super().__init__(res, d, METRIC_INNER_PRODUCT, device, use_tensor_cores, use_float16)


class GpuIndexIVF(Index):
    """faiss::gpu::GpuIndexIVF (faiss/gpu/GpuIndexIVF.h:40-167)."""

    @property
    def nprobe(self):
        return lib.faiss_GpuIndexIVF_nprobe(self._h)

    @nprobe.setter
    def nprobe(self, v):
        check(lib.faiss_GpuIndexIVF_set_nprobe(self._h, ctypes.c_size_t(int(v))))

    @property
    def nlist(self):
        return lib.faiss_GpuIndexIVF_nlist(self._h)

    def setClustering(self, niter=-1, seed=-1, max_points_per_centroid=-1):
        check(lib.faiss_GpuIndexIVF_set_clustering(self._h, int(niter), int(seed), int(max_points_per_centroid)))

    def reserveMemory(self, n):
        check(lib.faiss_GpuIndexIVF_reserveMemory(self._h, ctypes.c_size_t(int(n))))

    def reclaimMemory(self):
        out = ctypes.c_size_t()
        check(lib.faiss_GpuIndexIVF_reclaimMemory(self._h, ctypes.byref(out)))
        return out.value

    def getListLength(self, l):
        return lib.faiss_GpuIndexIVF_get_list_size(self._h, ctypes.c_size_t(int(l)))

    def _code_size(self):
        raise NotImplementedError

    def getListVectorData(self, l):
        n = self.getListLength(l)
        out = np.empty(n * self._code_size(), dtype=np.uint8)
        check(lib.faiss_GpuIndexIVF_getListVectorData(self._h, ctypes.c_size_t(int(l)), _ptr(out, _c_u8)))
        return out

    def getListIndices(self, l):
        n = self.getListLength(l)
        out = np.empty(n, dtype=np.int64)
        check(lib.faiss_GpuIndexIVF_getListIndices(self._h, ctypes.c_size_t(int(l)), _ptr(out, _c_i64)))
        return out

    def setCoarseCentroids(self, c):
        c = _as_f32(c)
        assert tuple(c.shape) == (self.nlist, self.d)
        check(lib.faiss_GpuIndexIVF_setCoarseCentroids(self._h, _ptr(c, _c_f)))

    def getCoarseCentroids(self):
        out = np.empty((self.nlist, self.d), dtype=np.float32)
        check(lib.faiss_GpuIndexIVF_getCoarseCentroids(self._h, _ptr(out, _c_f)))
        return out

    def setList(self, l, codes, ids):
        codes = np.ascontiguousarray(codes, dtype=np.uint8).reshape(-1)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        assert codes.size == ids.size * self._code_size()
        check(
            lib.faiss_GpuIndexIVF_setList(
                self._h, ctypes.c_size_t(int(l)), ctypes.c_int64(ids.size), _ptr(codes, _c_u8), _ptr(ids, _c_i64)
            )
        )

    def setListSizes(self, lens):
        """exact capacity for every list in one relayout (call before the per-list setList of a bulk clone)"""
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        assert lens.shape == (self.nlist,)
        check(lib.faiss_GpuIndexIVF_setListSizes(self._h, _ptr(lens, _c_i64)))

    def setIsTrained(self, v=True):
        check(lib.faiss_GpuIndexIVF_set_is_trained(self._h, int(bool(v))))

    def search_preassigned(self, x, k, assign, centroid_dis=None):
        x = self._check_x(x)
        n = x.shape[0]
        assign = _as_i64(assign)
        if centroid_dis is None:
            centroid_dis = np.zeros(assign.shape, dtype=np.float32)
            if _is_torch(x) and x.is_cuda:
                import torch

                centroid_dis = torch.zeros(tuple(assign.shape), dtype=torch.float32, device=x.device)
        centroid_dis = _as_f32(centroid_dis)
        D = _empty_like_residency(x, (n, k), np.float32)
        I = _empty_like_residency(x, (n, k), np.int64)
        self._use_torch_stream(x, assign)
        check(
            lib.faiss_GpuIndexIVF_search_preassigned(
                self._h,
                ctypes.c_int64(n),
                _ptr(x, _c_f),
                ctypes.c_int64(k),
                _ptr(assign, _c_i64),
                _ptr(centroid_dis, _c_f),
                _ptr(D, _c_f),
                _ptr(I, _c_i64),
            )
        )
        return D, I


class GpuIndexIVFFlat(GpuIndexIVF):
    """faiss::gpu::GpuIndexIVFFlat (faiss/gpu/GpuIndexIVFFlat.h:24-119)."""

    def __init__(self, res, d, nlist, metric=METRIC_L2, device=0, quantizer=None):
        """quantizer: a GpuIndexFlat to share as the coarse quantiser (faiss/gpu/GpuIndexIVFFlat.h:48-59), or None"""
        super().__init__()
        self._keep.append(res)
        if quantizer is not None:
            self._keep.append(quantizer)
            check(
                lib.faiss_GpuIndexIVFFlat_new_with_quantizer(
                    ctypes.byref(self._h), res._h, quantizer._h, int(d), ctypes.c_int64(nlist), int(metric), int(device)
                )
            )
            return
        check(
            lib.faiss_GpuIndexIVFFlat_new(
                ctypes.byref(self._h), res._h, int(d), ctypes.c_int64(nlist), int(metric), int(device)
            )
        )

    def _code_size(self):
        return 4 * self.d


class GpuIndexIVFPQ(GpuIndexIVF):
    """faiss::gpu::GpuIndexIVFPQ (faiss/gpu/GpuIndexIVFPQ.h:56-181)."""

    def __init__(self, res, d, nlist, M, nbits=8, metric=METRIC_L2, device=0, quantizer=None):
        """quantizer: a GpuIndexFlat to share as the coarse quantiser (faiss/gpu/GpuIndexIVFPQ.h:69-82), or None"""
        super().__init__()
        self._keep.append(res)
        self.M = int(M)
        if quantizer is not None:
            self._keep.append(quantizer)
            check(
                lib.faiss_GpuIndexIVFPQ_new_with_quantizer(
                    ctypes.byref(self._h), res._h, quantizer._h, int(d), ctypes.c_int64(nlist), ctypes.c_int64(M), ctypes.c_int64(nbits),
                    int(metric), int(device),
                )
            )
            return
        check(
            lib.faiss_GpuIndexIVFPQ_new(
                ctypes.byref(self._h),
                res._h,
                int(d),
                ctypes.c_int64(nlist),
                ctypes.c_int64(M),
                ctypes.c_int64(nbits),
                int(metric),
                int(device),
            )
        )

    def _code_size(self):
        return self.M

    def setPQCentroids(self, c):
        c = _as_f32(c)
        assert c.size == 256 * self.d
        check(lib.faiss_GpuIndexIVFPQ_setPQCentroids(self._h, _ptr(c, _c_f)))

    def getPQCentroids(self):
        out = np.empty((self.M, 256, self.d // self.M), dtype=np.float32)
        check(lib.faiss_GpuIndexIVFPQ_getPQCentroids(self._h, _ptr(out, _c_f)))
        return out

    def setPQClustering(self, niter=-1, seed=-1, max_points_per_centroid=-1):
        check(lib.faiss_GpuIndexIVFPQ_set_pq_clustering(self._h, int(niter), int(seed), int(max_points_per_centroid)))

    def setPrecomputedCodes(self, enable):
        check(lib.faiss_GpuIndexIVFPQ_setPrecomputedCodes(self._h, int(bool(enable))))


class IndexShards(Index):
    """faiss::IndexShards (faiss/IndexShards.h:21-106)."""

    def __init__(self, d, threaded=False, successive_ids=True):
        super().__init__()
        check(
            lib.faiss_IndexShards_new_with_options(
                ctypes.byref(self._h), ctypes.c_int64(d), int(bool(threaded)), int(bool(successive_ids))
            )
        )
        self._shards = []

    def add_shard(self, index):
        check(lib.faiss_IndexShards_add_shard(self._h, index._h))
        self._shards.append(index)

    def _resources(self):
        out = []
        for s in self._shards:
            out.extend(s._resources())
        return out

    def remove_shard(self, index):
        check(lib.faiss_IndexShards_remove_shard(self._h, index._h))
        self._shards.remove(index)

    def at(self, i):
        return self._shards[i]

    def count(self):
        return len(self._shards)

    def lastSearchPath(self):
        """'nccl' if the last search ran per-device threads + ncclAllGather + device merge, 'host' for the
        reference's thread-per-shard + host merge"""
        return {0: "host", 1: "nccl"}.get(lib.faiss_IndexShards_lastSearchPath(self._h), "?")

    def __del__(self):
        # free the meta index before the shards it points to
        if getattr(self, "_h", None) and lib is not None:
            lib.faiss_Index_free(self._h)
            self._h = None
        self._shards = []


class IndexShardsIVF(IndexShards):
    """faiss::IndexShardsIVF (faiss/IndexShardsIVF.cpp:100-251): IVF shards over one shared coarse quantiser (a
    GpuIndexFlat): one coarse search, search_preassigned on every shard, merge."""

    def __init__(self, quantizer, nlist, threaded=False, successive_ids=True):
        Index.__init__(self)
        self._keep.append(quantizer)
        check(lib.faiss_IndexShardsIVF_new(ctypes.byref(self._h), quantizer._h, ctypes.c_int64(nlist), int(bool(threaded)), int(bool(successive_ids))))
        self._shards = []

    def add_shard(self, index):
        check(lib.faiss_IndexShardsIVF_add_shard(self._h, index._h))
        self._shards.append(index)


def nccl_unique_id():
    """128 bytes to hand to every rank's StandardGpuResources.ncclInitRank (ncclGetUniqueId)."""
    buf = ctypes.create_string_buffer(128)
    check(lib.faiss_b200_nccl_unique_id(buf))
    return buf.raw


class DistributedIndexShards(Index):
    """IndexShards with one shard per NCCL rank (faiss/IndexShards.cpp:197-264 semantics; one grouped
    all-gather + device merge; Flat shards pool their thresholds).  `search` is a collective call."""

    def __init__(self, res, local, successive_ids=True):
        super().__init__()
        self._keep.append(res)
        self._local = local
        check(lib.faiss_DistributedIndexShards_new(ctypes.byref(self._h), res._h, local._h, int(bool(successive_ids))))

    def _resources(self):
        return [r for r in self._keep if isinstance(r, StandardGpuResources)]

    def sync(self):
        check(lib.faiss_DistributedIndexShards_sync(self._h))

    def info(self):
        r, n, o = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
        check(lib.faiss_DistributedIndexShards_info(self._h, ctypes.byref(r), ctypes.byref(n), ctypes.byref(o)))
        return {"rank": r.value, "world": n.value, "id_offset": o.value}

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.faiss_Index_free(self._h)
            self._h = None
        self._local = None


def kmeans(res, x, k, niter=25, seed=1234, max_points_per_centroid=256, device=0):
    """Lloyd k-means on the device (role of faiss.Kmeans / faiss::Clustering); returns (centroids, obj)."""
    x = _as_f32(x)
    n, d = x.shape
    cent = np.empty((k, d), dtype=np.float32)
    obj = np.zeros(niter, dtype=np.float32)
    check(
        lib.faiss_b200_kmeans(
            res._h,
            int(device),
            ctypes.c_size_t(d),
            ctypes.c_size_t(n),
            ctypes.c_size_t(k),
            _ptr(x, _c_f),
            int(niter),
            int(seed),
            int(max_points_per_centroid),
            _ptr(cent, _c_f),
            _ptr(obj, _c_f),
        )
    )
    return cent, obj


def kmeans_ex(res, x, k, niter=25, seed=1234, max_points_per_centroid=256, metric=METRIC_L2, spherical=False, device=0):
    """k-means with an assignment index of `metric` and ClusteringParameters::spherical."""
    x = _as_f32(x)
    n, d = x.shape
    cent = np.empty((k, d), dtype=np.float32)
    obj = np.zeros(niter, dtype=np.float32)
    check(
        lib.faiss_b200_kmeans_ex(
            res._h, int(device), ctypes.c_size_t(d), ctypes.c_size_t(n), ctypes.c_size_t(k), _ptr(x, _c_f), int(niter), int(seed),
            int(max_points_per_centroid), int(metric), int(bool(spherical)), _ptr(cent, _c_f), _ptr(obj, _c_f),
        )
    )
    return cent, obj


def bfKnn(res, xq, xb, k, metric=METRIC_L2, device=0):
    """faiss.knn_gpu / bfKnn (faiss/gpu/GpuDistance.h:33-181, faiss/python/gpu_wrappers.py:60-200): brute-force
    k-NN of xq in xb, row-major fp32, numpy or torch CUDA inputs; outputs follow xq's residency."""
    xq, xb = _as_f32(xq), _as_f32(xb)
    nq, d = xq.shape
    assert xb.shape[1] == d
    D = _empty_like_residency(xq, (nq, k), np.float32)
    I = _empty_like_residency(xq, (nq, k), np.int64)
    if _is_torch(xq) and xq.is_cuda:
        import torch

        res.setDefaultStream(device, torch.cuda.current_stream(device).cuda_stream)
    check(
        lib.faiss_b200_bfKnn(
            res._h, int(device), int(metric), ctypes.c_int64(k), int(d), _ptr(xb, _c_f), ctypes.c_int64(xb.shape[0]), _ptr(xq, _c_f),
            ctypes.c_int64(nq), _ptr(D, _c_f), _ptr(I, _c_i64),
        )
    )
    return D, I


def knn_gpu(res, xq, xb, k, D=None, I=None, metric=METRIC_L2, device=0):
    """faiss.knn_gpu (faiss/python/gpu_wrappers.py:60-200) for row-major fp32 inputs: argument order of the reference,
    optional preallocated outputs."""
    rD, rI = bfKnn(res, xq, xb, k, metric, device)
    if D is not None:
        D[...] = rD
        rD = D
    if I is not None:
        I[...] = rI
        rI = I
    return rD, rI


def kmeans_sharded(res, x_local, k, niter=25, seed=1234, device=0):
    """Collective: k-means over the rows of ALL ranks of the device's NCCL communicator (this rank passes its own
    rows; rank order = row order).  Returns (centroids [k, d] identical on every rank, objective per iteration,
    stats dict)."""
    x_local = _as_f32(x_local)
    n, d = x_local.shape
    cent = np.empty((k, d), dtype=np.float32)
    obj = np.zeros(niter, dtype=np.float32)
    st = (ctypes.c_double * 4)()
    check(
        lib.faiss_b200_kmeans_sharded(
            res._h, int(device), ctypes.c_size_t(d), ctypes.c_size_t(n), ctypes.c_size_t(k), _ptr(x_local, _c_f), int(niter), int(seed),
            _ptr(cent, _c_f), _ptr(obj, _c_f), st,
        )
    )
    return cent, obj, {"total_s": st[0], "search_update_allreduce_s": st[1], "split_clusters_s": st[2], "nsplit": int(st[3])}


def pq_train(res, x, M, niter=25, seed=1234, device=0):
    """faiss::ProductQuantizer::train (M independent 256-centroid k-means); returns [M, 256, d/M]."""
    x = _as_f32(x)
    n, d = x.shape
    out = np.empty((M, 256, d // M), dtype=np.float32)
    check(
        lib.faiss_b200_pq_train(
            res._h, int(device), ctypes.c_size_t(d), ctypes.c_size_t(M), ctypes.c_size_t(n), _ptr(x, _c_f), int(niter), int(seed),
            _ptr(out, _c_f),
        )
    )
    return out


# ------------------------------------------------------------------ tier-2 seams (torch CUDA tensors)
def flat_search_exact(res, Y, Q, k, metric=METRIC_L2, device=0):
    import torch

    assert Y.is_cuda and Q.is_cuda
    res.setDefaultStream(device, torch.cuda.current_stream(device).cuda_stream)
    nq = Q.shape[0]
    D = torch.empty((nq, k), dtype=torch.float32, device=Q.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=Q.device)
    check(
        lib.b200_flat_search_exact(
            res._h, int(device), _ptr(Y, _c_f), ctypes.c_int64(Y.shape[0]), int(Y.shape[1]), _ptr(Q, _c_f),
            ctypes.c_int64(nq), int(k), int(metric), _ptr(D, _c_f), _ptr(I, _c_i64),
        )
    )
    return D, I


def kmeans_accumulate(res, x, assign, k, device=0):
    """Per-centroid partial sums [k, d] and counts [k] of the CUDA rows `x` under `assign` (int64 [n]) --
    the device half of compute_centroids (faiss/impl/ClusteringHelpers.cpp:101-172); the caller
    all-reduces them across shards and divides."""
    import torch

    assert x.is_cuda and assign.is_cuda
    res.setDefaultStream(device, torch.cuda.current_stream(device).cuda_stream)
    n, d = x.shape
    sums = torch.zeros((k, d), dtype=torch.float32, device=x.device)
    counts = torch.zeros((k,), dtype=torch.float32, device=x.device)
    check(
        lib.b200_kmeans_update(
            res._h, int(device), _ptr(x, _c_f), _ptr(assign, _c_i64), ctypes.c_int64(n), int(d), ctypes.c_int64(k),
            _ptr(sums, _c_f), _ptr(counts, _c_f), None,
        )
    )
    return sums, counts


def rand_perm(n, seed):
    """faiss::rand_perm (faiss/utils/random.cpp:130-142), host."""
    perm = np.empty(int(n), dtype=np.int32)
    check(lib.faiss_b200_rand_perm(perm.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), ctypes.c_size_t(int(n)), ctypes.c_int64(int(seed))))
    return perm


def split_clusters(hassign, centroids, n):
    """faiss split_clusters (faiss/impl/ClusteringHelpers.cpp:177-240) in place on host arrays; returns nsplit."""
    k, d = centroids.shape
    assert hassign.dtype == np.float32 and centroids.dtype == np.float32
    ns = ctypes.c_int(0)
    fp = ctypes.POINTER(ctypes.c_float)
    check(
        lib.faiss_b200_split_clusters(
            ctypes.c_size_t(d), ctypes.c_size_t(k), ctypes.c_size_t(int(n)), hassign.ctypes.data_as(fp), centroids.ctypes.data_as(fp),
            ctypes.byref(ns),
        )
    )
    return ns.value


def topk_merge(res, D_in, I_in, k, metric=METRIC_L2, id_offsets=None, device=0):
    """D_in/I_in: CUDA tensors [nq, nshard, k_in] -> merged [nq, k] (role of merge_knn_results)."""
    import torch

    nq, nshard, kin = D_in.shape
    res.setDefaultStream(device, torch.cuda.current_stream(device).cuda_stream)
    D = torch.empty((nq, k), dtype=torch.float32, device=D_in.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=D_in.device)
    check(
        lib.b200_topk_merge(
            res._h, int(device), _ptr(D_in, _c_f), _ptr(I_in, _c_i64), ctypes.c_int64(nq), int(nshard), int(kin),
            _ptr(id_offsets, _c_i64), int(k), int(metric), _ptr(D, _c_f), _ptr(I, _c_i64),
        )
    )
    return D, I


def flat_tc_scores_debug(res, Q16, Y16, device=0):
    """Raw tcgen05 fp16 score matrix [nq, roundup(N,128)] (unit-test seam)."""
    import torch

    nq, dpad = Q16.shape
    N = Y16.shape[0]
    res.setDefaultStream(device, torch.cuda.current_stream(device).cuda_stream)
    npad = (N + 255) // 256 * 256
    S = torch.zeros((nq, npad), dtype=torch.float32, device=Q16.device)
    check(
        lib.b200_flat_tc_scores_debug(
            res._h, int(device), ctypes.c_void_p(Q16.data_ptr()), ctypes.c_int64(nq), ctypes.c_void_p(Y16.data_ptr()),
            ctypes.c_int64(N), int(dpad), _ptr(S, _c_f),
        )
    )
    return S

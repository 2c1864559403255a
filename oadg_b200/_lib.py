"""ctypes binding of libOADG.so (C ABI declared in include/oadg.h).

The product path has no CPU fallback: if the CUDA library cannot be loaded, or no
CUDA device is present when a compute entry point is called, this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('OADG_LIB') or os.path.join(_HERE, 'libOADG.so')   # OADG_LIB: a timing variant (scripts/)

_c = ctypes
_vp, _i32, _sz, _f32 = _c.c_void_p, _c.c_int32, _c.c_size_t, _c.c_float

# name -> (restype, argtypes); must list every symbol include/oadg.h declares
SIGNATURES = {
    'oadg_abi_version': (_c.c_int, []),
    'oadg_struct_sizes': (None, [_c.POINTER(_i32)]),
    'oadg_error_string': (_c.c_char_p, [_c.c_int]),
    'oadg_saliency_scores': (_c.c_int, [_vp, _vp, _vp, _c.c_int, _vp, _vp]),
    'oadg_oamix_workspace_bytes': (_c.c_int, [_vp, _sz, _c.POINTER(_sz)]),
    'oadg_memcpy_async': (_c.c_int, [_vp, _vp, _sz, _c.c_int, _vp]),
    'oadg_oamix_execute': (_c.c_int, [_vp, _sz, _vp, _c.c_int, _vp, _vp, _sz, _c.POINTER(_c.c_int), _vp]),
    'oadg_oamix_execute_fused': (_c.c_int, [_vp, _sz, _vp, _c.c_int, _vp, _vp, _vp, _sz, _c.POINTER(_c.c_int), _vp]),
    'oadg_oamix_execute_shared': (_c.c_int, [_vp, _sz, _vp, _c.c_int, _vp, _vp, _sz, _c.c_int, _c.POINTER(_c.c_int),
                                             _vp]),
    'oadg_oamix_execute_profiled': (_c.c_int, [_vp, _sz, _vp, _c.c_int, _vp, _vp, _sz, _c.POINTER(_f32),
                                               _c.POINTER(_f32), _c.POINTER(_c.c_int), _c.POINTER(_c.c_int), _vp, _vp]),
    'oadg_oamix_last_trace': (_c.c_int, [_vp, _c.c_int]),
    'oadg_oamix_poll_fault': (_c.c_int, [_c.c_int]),
    'oadg_oamix_sample_plan': (_c.c_int, [_vp, _vp, _c.c_int, _vp, _vp, _vp, _vp, _vp, _sz, _c.POINTER(_sz),
                                          _vp, _vp, _vp, _vp, _vp]),
    'oadg_supcon_workspace_bytes': (_c.c_int, [_c.c_int, _c.c_int, _c.POINTER(_sz)]),
    'oadg_supcon_forward': (_c.c_int, [_vp, _vp, _vp, _c.c_int, _c.c_int, _f32, _f32, _c.c_int, _c.c_int,
                                       _vp, _vp, _sz, _c.POINTER(_c.c_int), _vp]),
    'oadg_supcon_backward': (_c.c_int, [_vp, _vp, _vp, _c.c_int, _c.c_int, _f32, _f32, _c.c_int,
                                        _vp, _vp, _vp, _sz, _c.POINTER(_c.c_int), _vp]),
    'oadg_supcon_normalize': (_c.c_int, [_vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _vp, _vp, _sz, _vp]),
    'oadg_supcon_forward_gathered': (_c.c_int, [_vp, _vp, _vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _f32, _f32,
                                                _c.c_int, _vp, _vp, _vp, _sz, _c.POINTER(_c.c_int), _vp]),
    'oadg_supcon_backward_gathered': (_c.c_int, [_vp, _vp, _vp, _vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _f32,
                                                 _c.c_int, _vp, _vp, _vp, _sz, _c.POINTER(_c.c_int), _vp]),
    'oadg_supcon_pack_width': (_c.c_int, [_c.c_int]),
    'oadg_supcon_gather_pack': (_c.c_int, [_vp, _vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _vp, _vp, _sz, _vp]),
    'oadg_supcon_forward_packed': (_c.c_int, [_vp, _vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _f32, _f32, _c.c_int,
                                              _vp, _vp, _sz, _c.POINTER(_c.c_int), _vp]),
    'oadg_supcon_finish_packed': (_c.c_int, [_vp, _c.c_int, _c.c_int, _c.c_int, _vp, _vp, _sz, _vp]),
    'oadg_supcon_backward_packed': (_c.c_int, [_vp, _vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _f32, _c.c_int, _vp,
                                               _vp, _vp, _sz, _c.POINTER(_c.c_int), _vp]),
    'oadg_peer_alloc': (_c.c_int, [_sz, _c.POINTER(_vp)]),
    'oadg_peer_free': (_c.c_int, [_vp]),
    'oadg_peer_export': (_c.c_int, [_vp, _vp]),
    'oadg_peer_import': (_c.c_int, [_vp, _c.POINTER(_vp)]),
    'oadg_peer_release': (_c.c_int, [_vp]),
    'oadg_peer_fault_alloc': (_c.c_int, [_c.POINTER(_c.POINTER(_c.c_uint32))]),
    'oadg_supcon_gather_pack_peers': (_c.c_int, [_vp, _vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _vp, _sz, _sz,
                                                 _sz, _c.c_uint32, _vp, _sz, _vp]),
    'oadg_peer_scatter': (_c.c_int, [_vp, _sz, _sz, _sz, _sz, _c.c_uint32, _c.c_uint32, _vp]),
    'oadg_peer_wait': (_c.c_int, [_vp, _c.c_int, _c.c_uint32, _c.c_uint32, _c.c_uint32, _vp, _vp]),
    'oadg_jsd2_scratch_bytes': (_c.c_int, []),
    'oadg_jsd2_forward': (_c.c_int, [_vp, _c.c_int, _c.c_int, _vp, _vp, _vp, _vp]),
}

_lib = None


class OADGError(RuntimeError):
    pass


def load():
    """Load libOADG.so (built in-tree by ``python -m oadg_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OADGError(
            'libOADG.so not found at %s: build it with `python -m oadg_b200.build` '
            '(nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.oadg_abi_version() != 1:
        raise OADGError('libOADG.so ABI mismatch')
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = load().oadg_error_string(code)
        raise OADGError('libOADG error %d: %s' % (code, msg.decode() if msg else '?'))


_torch_cuda = None


def require_cuda():
    """torch, after checking once that a CUDA device is there (the answer does not change within a process)."""
    global _torch_cuda
    if _torch_cuda is None:
        import torch
        if not torch.cuda.is_available():
            raise OADGError('oadg_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback')
        _torch_cuda = torch
    return _torch_cuda


def raw_stream(device):
    """cudaStream_t (int) of torch's current stream on ``device``: the C accessor when this torch has it (a few
    hundred ns instead of building a torch.cuda.Stream object), torch.cuda.current_stream otherwise."""
    import torch
    idx = device.index if device.index is not None else torch.cuda.current_device()
    try:
        return torch._C._cuda_getCurrentRawStream(idx)
    except AttributeError:
        return torch.cuda.current_stream(device).cuda_stream

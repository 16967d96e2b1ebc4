"""Device plumbing shared by the harness modules: torch owns memory and streams, nothing else."""
import ctypes as C
import numpy as np
import torch


def stream():
    """Handle of torch's current CUDA stream (all dgb_* calls are enqueued on it)."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device address of a float64/int32/int64 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "feltor_b200 operates on contiguous CUDA tensors only"
    return C.c_void_p(t.data_ptr())


def dvec(a):
    """Host numpy array -> device tensor (dtype preserved)."""
    a = np.ascontiguousarray(a)
    if not a.flags.writeable:  # e.g. a broadcast view: torch refuses to wrap read-only memory silently
        a = a.copy()
    return torch.from_numpy(a).cuda()


def hvec(t):
    return t.detach().cpu().numpy()


def hptr(a):
    """Address of a C-contiguous numpy array."""
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)

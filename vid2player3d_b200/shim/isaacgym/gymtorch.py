"""`isaacgym.gymtorch` shim: wrap/unwrap are identity (state tensors are torch-owned;
the CUDA extension receives raw data_ptr()s through include/b200env.h)."""


def wrap_tensor(t, *a, **k):
    return t


def unwrap_tensor(t, *a, **k):
    return t

"""Host-side mirror of the reference's `mmgpt` package for the Merlin hot path.

Module paths, class names, constructor/forward signatures and error behaviour follow
/root/reference/mmgpt so that reference call sites keep working; the arithmetic
runs in libmerlin_b200.so (hand-written sm_100a kernels) through merlin_b200.ops.
"""

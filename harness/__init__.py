"""Shims that reuse the reference's harness-side files UNMODIFIED, in place, when
/root/reference is mounted (dev container); nothing here is needed at run time on a GPU box."""

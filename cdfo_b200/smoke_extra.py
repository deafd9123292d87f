"""Additional smoke checks appended as kernels land (called by __graft_entry__.smoke)."""


def run(dev):
    return None

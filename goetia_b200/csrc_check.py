"""Host restatement of the device fastmod (kernels.cuh fastmod_u64), used by tests to check the
arithmetic for divisors too large to allocate a table for."""
_M = (1 << 64) - 1


def fastmod_host(h, d):
    m = _M // d
    q = (h * m) >> 64
    r = (h - q * d) & _M
    return r - d if r >= d else r

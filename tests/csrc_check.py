"""Host restatement of the device fastmod (kernels.cuh fastmod_u64), used by tests to check the
arithmetic for divisors too large to allocate a table for."""
_M = (1 << 64) - 1


def fastmod_host(h, d):
    m = _M // d
    q = (h * m) >> 64
    r = (h - q * d) & _M
    return r - d if r >= d else r


def bin_of_host(h, d):
    """Host restatement of the device bin_of (bucket.cuh): the 32-bit-reciprocal reduction used for
    tables of >= 2^28 slots.  Returns None where the device falls back to fastmod_u64."""
    rs = 0
    while rs <= 4 and (d << rs) < (1 << 32):
        rs += 1
    if rs > 4 or d > (1 << 58):
        return None
    dsh = d << rs
    m32 = _M // dsh
    assert m32 < (1 << 32)
    hl, hh = h & 0xFFFFFFFF, h >> 32
    u = hh * m32 + ((hl * m32) >> 32)
    assert u < (1 << 64)
    q = u >> 32
    qd = (q * (dsh & 0xFFFFFFFF) + (((q * (dsh >> 32)) & 0xFFFFFFFF) << 32)) & _M
    r = (h - qd) & _M
    if r >= dsh:
        r -= dsh
    for j in range(3, -1, -1):
        if j < rs and r >= (d << j):
            r -= d << j
    return r

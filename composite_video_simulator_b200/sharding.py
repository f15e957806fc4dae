"""Field sharding across ranks (one process per GPU).

Fields are independent work items once the rand() position is known, and that position is a
closed form of the field index: draws(field f) depends only on the field's
parity, so the stream position of global field index g is

    pos(g) = ceil(g/2) * draws(parity of field 0) + floor(g/2) * draws(parity of field 1)

Every step the job processes world * batch consecutive fields; rank r takes the r-th contiguous
chunk.  No pixel ever crosses ranks: there is no data-path collective.
"""
from . import api


def field_parity(fieldno):
    """The reference loop's schedule, ffmpeg_ntsc.cpp:2229: field = (current & 1) ^ 1."""
    return (fieldno & 1) ^ 1


def stream_position(params, w, h, g):
    """rand() draws consumed before global field index g (fields 0..g-1 processed in order)."""
    d_even = api.draws_per_field(params, w, h, field_parity(0))    # fields 0, 2, 4, ...
    d_odd = api.draws_per_field(params, w, h, field_parity(1))     # fields 1, 3, 5, ...
    return ((g + 1) // 2) * d_even + (g // 2) * d_odd


def chunk(step, rank, world, batch):
    """(first global field index, count) of rank's share of one step."""
    return (step * world + rank) * batch, batch


def stream_position_yuv422(params, w, h, g):
    """The same closed form for the 4:2:2 path (include/cvs_yuv422.h): its draw layout differs
    (ffmpeg_to_composite.cpp:653-941) but it also depends on the field parity only."""
    import ctypes as C
    from . import yuv422
    L = yuv422._bind(["cvs422_draws_per_field"])
    d_even = L.cvs422_draws_per_field(C.byref(params), w, h, field_parity(0))
    d_odd = L.cvs422_draws_per_field(C.byref(params), w, h, field_parity(1))
    return ((g + 1) // 2) * d_even + (g // 2) * d_odd

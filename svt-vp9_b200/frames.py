"""Host-side preparation of the planes motion estimation reads, mirroring the reference's picture-analysis
front end: edge-replicated padding of the source luma (EbEncHandle.c:1003-1006: 64+4 samples per side) and the
1/4 and 1/16 point decimations with 32 / 16 samples of padding (eb_vp9_decimation_2d + eb_vp9_generate_padding,
Source/Lib/Codec/EbPictureAnalysisProcess.c:102-122, :5025-5085)."""
import numpy as np

from . import api

FULL_PAD, QUARTER_PAD, SIXTEENTH_PAD = 68, 32, 16


def pa_planes(luma):
    """luma: (H, W) uint8 -> dict of padded C-contiguous arrays 'full', 'quarter', 'sixteenth'."""
    assert luma.dtype == np.uint8 and luma.ndim == 2
    full = np.ascontiguousarray(np.pad(luma, FULL_PAD, mode="edge"))
    quarter = np.ascontiguousarray(np.pad(luma[::2, ::2], QUARTER_PAD, mode="edge"))
    sixteenth = np.ascontiguousarray(np.pad(luma[::4, ::4], SIXTEENTH_PAD, mode="edge"))
    return {"full": full, "quarter": quarter, "sixteenth": sixteenth, "w": luma.shape[1], "h": luma.shape[0]}


def plane_structs(p):
    """(api.Plane * 3) describing the arrays of pa_planes() (the arrays must outlive the structs)."""
    w, h = p["w"], p["h"]
    arr = (api.Plane * 3)()
    arr[0] = api.plane_of(p["full"], FULL_PAD, FULL_PAD, w, h)
    arr[1] = api.plane_of(p["quarter"], QUARTER_PAD, QUARTER_PAD, (w + 1) // 2, (h + 1) // 2)
    arr[2] = api.plane_of(p["sixteenth"], SIXTEENTH_PAD, SIXTEENTH_PAD, (w + 3) // 4, (h + 3) // 4)
    return arr
